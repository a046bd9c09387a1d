"""CPU, build container only: the C restatement against the reference's OWN sources compiled
into oracle/_ref (skipped where /root/reference and a prebuilt _ref are both absent)."""
import numpy as np
import pytest

from tests import oracle_lib


@pytest.mark.parametrize("depth,seed", [(3, 12345), (4, 1), (5, 5), (6, 12345), (7, 70), (8, 12345), (8, 777), (9, 9), (10, 12345)])
@pytest.mark.parametrize("volcano", [0, 1])
def test_restatement_equals_reference_bitwise(oracle, ref, depth, seed, volcano):
    r, calls = ref.terrain2d(depth, seed, volcano=volcano)
    o = oracle.terrain2d(oracle_lib.params(seed=seed, depth=depth, apply_volcano=volcano))
    assert calls == oracle.L.rno_draw_count(depth)
    assert np.array_equal(r.view(np.uint32), o.view(np.uint32))


def test_reference_amplitude_argument(oracle, ref):
    r, _ = ref.terrain2d(7, 42, amplitude=0.35, volcano=1)
    o = oracle.terrain2d(oracle_lib.params(seed=42, depth=7, apply_volcano=1, amplitude=0.35))
    assert np.array_equal(r.view(np.uint32), o.view(np.uint32))


def test_reference_random_function(oracle, ref):
    """basics.h:16-19 evaluated by the reference's own inline random()."""
    for state in (0, 1, 12345, 0xDEADBEEF, 0xFFFFFFFF):
        nxt = oracle.L.rno_lcg_next_state(state)
        assert ref.L.ref_random_from_state(state) == oracle.L.rno_random_from_draw((nxt >> 16) & 0x7FFF)


def test_reference_lcg(oracle, ref):
    for seed in (1, 12345, 0xFFFFFFFF):
        d = ref.lcg_draws(seed, 64)
        assert [oracle.L.rno_lcg_draw(seed, k) for k in range(64)] == d.tolist()


@pytest.mark.parametrize("depth,seed", [(3, 12345), (4, 9), (5, 7), (6, 1), (8, 12345), (9, 31337), (10, 777)])
def test_cell_quadtree_restatement_equals_reference(oracle, ref, depth, seed):
    """Terrain::generateRecursive (Terrain.cpp:260-356) run by the reference itself vs rn_oracle_cells.c: positions,
    colours, radii, quantised normals and the per-level visiting order, byte for byte."""
    field, rc = ref.cells(depth, seed)
    p = oracle_lib.params(seed=seed, depth=depth, apply_volcano=1)
    f2 = oracle.terrain2d(p)
    assert np.array_equal(field.view(np.uint32), f2.view(np.uint32))
    oc = oracle.cells(p, f2)
    assert rc.tobytes() == oc.tobytes()
