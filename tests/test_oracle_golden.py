"""CPU: the C restatement (oracle/rn_oracle.c) against the committed golden vectors that were
produced by the reference's own sources (tests/golden/make_golden.py), plus its internal
consistency (sequential vs index-addressed form)."""
import json
import os

import numpy as np
import pytest

from tests import oracle_lib

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "terrain2d_golden.json")))


def test_lcg_known_answers(oracle):
    # MSVC CRT known answers for srand(1) and the reference-build KATs
    assert GOLD["lcg_kat"]["1"][:8] == [41, 18467, 6334, 26500, 19169, 15724, 11478, 29358]
    for seed_s, draws in GOLD["lcg_kat"].items():
        seed = int(seed_s)
        assert [oracle.L.rno_lcg_draw(seed, k) for k in range(len(draws))] == draws


def test_lcg_skip_ahead_matches_sequential(oracle):
    rng = np.random.default_rng(0)
    for seed in (0, 1, 12345, 0xFFFFFFFF, 0x80000000):
        s = seed
        for k in range(1, 300):
            s = oracle.L.rno_lcg_next_state(s)
            assert oracle.L.rno_lcg_state_after(seed, k) == s
    for _ in range(200):
        seed = int(rng.integers(0, 2**32))
        k = int(rng.integers(0, 2**40))
        a = oracle.L.rno_lcg_state_after(seed, k)
        b = oracle.L.rno_lcg_state_after(oracle.L.rno_lcg_state_after(seed, k // 3), k - k // 3)
        assert a == b
    # full period 2^32
    assert oracle.L.rno_lcg_state_after(777, 2**32) == 777


def test_random_from_draw_range(oracle):
    assert oracle.L.rno_random_from_draw(0) == -1.0
    assert oracle.L.rno_random_from_draw(0x7FFF) == 1.0
    v = [oracle.L.rno_random_from_draw(r) for r in range(0, 0x8000, 97)]
    assert all(-1.0 <= x <= 1.0 for x in v) and v == sorted(v)


@pytest.mark.parametrize("depth", [3, 4, 5, 6, 8, 10, 12])
def test_draw_count(oracle, depth):
    want = {e["depth"]: e["rand_calls"] for e in GOLD["fields"]}[depth]
    assert oracle.L.rno_draw_count(depth) == want


@pytest.mark.parametrize("depth", [3, 4, 5, 6])
def test_call_index_is_a_bijection(oracle, depth):
    W = 1 << depth
    seen = {}
    for i in range(W + 1):
        for j in range(W + 1):
            k = oracle.L.rno_call_index(depth, i, j)
            if i % (W // 4) == 0 and j % (W // 4) == 0:
                assert k == 2**64 - 1
            else:
                assert k not in seen
                seen[k] = (i, j)
    assert sorted(seen) == list(range(oracle.L.rno_draw_count(depth)))


@pytest.mark.parametrize("entry", GOLD["fields"], ids=lambda e: f"d{e['depth']}-s{e['seed']}")
def test_seq_matches_reference_digests(oracle, entry):
    for volc, key in ((0, "fnv64_pre_volcano"), (1, "fnv64_post_volcano")):
        p = oracle_lib.params(seed=entry["seed"], depth=entry["depth"], apply_volcano=volc)
        f = oracle.terrain2d(p)
        assert oracle.fnv64(f) == entry[key]
        if volc:
            y = f[..., 1].astype(np.float64)
            assert abs(y.mean() - entry["post_mean_y"]) < 1e-12
            assert y.min() == entry["post_min_y"] and y.max() == entry["post_max_y"]


@pytest.mark.parametrize("entry", GOLD["small_fields"], ids=lambda e: f"d{e['depth']}")
def test_small_fields_bitwise(oracle, entry):
    for volc, key in ((0, "pre_volcano_bits"), (1, "post_volcano_bits")):
        p = oracle_lib.params(seed=entry["seed"], depth=entry["depth"], apply_volcano=volc)
        got = oracle.terrain2d(p).view(np.uint32).reshape(-1)
        assert np.array_equal(got, np.array(entry[key], np.uint32))


def test_golden_rows(oracle):
    cache = {}
    for r in GOLD["rows"]:
        key = (r["depth"], r["seed"])
        if key not in cache:
            cache[key] = oracle.terrain2d(oracle_lib.params(seed=r["seed"], depth=r["depth"], apply_volcano=1))
        got = cache[key][r["row"]].view(np.uint32).reshape(-1)
        assert np.array_equal(got, np.array(r["bits"], np.uint32)), (key, r["row"])


@pytest.mark.parametrize("depth,seed", [(3, 12345), (4, 9), (6, 12345), (8, 777), (10, 12345)])
def test_indexed_form_equals_sequential(oracle, depth, seed):
    """Index-addressed RNG + ordered shade gather (what the kernels do), in forward and
    reversed visiting order, equals the sequential scatter form bit for bit."""
    p = oracle_lib.params(seed=seed, depth=depth, apply_volcano=0)
    a = oracle.terrain2d(p)
    assert np.array_equal(a.view(np.uint32), oracle.terrain2d_indexed(p, False).view(np.uint32))
    assert np.array_equal(a.view(np.uint32), oracle.terrain2d_indexed(p, True).view(np.uint32))


def test_indexed_nondefault_params(oracle):
    p = oracle_lib.params(seed=99, depth=6, apply_volcano=0, persistence=0.6, sideways_scale=0.3, radiosity_strength=0.7,
                          feature_scale=50.0, amplitude_variation=0.25, amplitude=0.35, bulgy_rocks=0)
    a = oracle.terrain2d(p)
    assert np.array_equal(a.view(np.uint32), oracle.terrain2d_indexed(p, True).view(np.uint32))
    q = oracle_lib.params(seed=99, depth=6, apply_volcano=0)
    assert not np.array_equal(a, oracle.terrain2d(q))


def test_spot_values_from_survey(oracle):
    """SURVEY.md App. B spot values (depth 8, seed 12345, post-volcano)."""
    f, vpos = oracle.terrain2d(oracle_lib.params(seed=12345, depth=8, apply_volcano=1), want_volcano_pos=True)
    np.testing.assert_array_equal(f[128, 128], np.array([0, 0.0416000038, 0, 1.00484598], np.float32))
    np.testing.assert_array_equal(f[1, 1], np.array([-0.992191017, -0.0400104895, -0.992338538, 1.03925121], np.float32))
    assert vpos == (0.0, float(np.float32(0.0416000038) * np.float32(500.0)), 0.0)


def test_bad_params(oracle):
    import ctypes as C
    buf = np.empty((9, 9, 4), np.float32)
    for kw in (dict(depth=2), dict(depth=15), dict(lacunarity=3), dict(persistence=0.0)):
        p = oracle_lib.params(**{**dict(depth=3), **kw})
        assert oracle.L.rno_terrain2d(C.byref(p), buf.ctypes.data, None) != 0


def test_batch_threads(oracle):
    p = oracle_lib.params(depth=6, apply_volcano=1)
    seeds = [12345, 1, 2, 3, 4]
    b = oracle.terrain2d_batch(p, seeds, threads=3)
    for t, s in enumerate(seeds):
        assert np.array_equal(b[t], oracle.terrain2d(oracle_lib.params(seed=s, depth=6, apply_volcano=1)))


@pytest.mark.parametrize("entry", GOLD.get("cells", []), ids=lambda e: f"d{e['depth']}-s{e['seed']}")
def test_cell_quadtree_matches_reference_digests(oracle, entry):
    """The Cell-quadtree restatement against digests produced by the reference's own generateRecursive."""
    p = oracle_lib.params(seed=entry["seed"], depth=entry["depth"], apply_volcano=1)
    cells = oracle.cells(p, oracle.terrain2d(p))
    assert len(cells) == entry["count"] == oracle_lib.cell_count(entry["depth"])
    assert oracle.fnv64(cells.view(np.uint8)) == entry["fnv64"]
    assert cells[:5].view(np.uint32).reshape(-1).tolist() == entry["first"]
