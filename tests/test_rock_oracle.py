"""CPU: the Rock restatement (oracle/rn_oracle_rock.c, Rock.cpp) against the reference's own functions (oracle/_ref, build
container only), against the committed golden digests produced by them, and against its own index-addressed form -- the
formulation the CUDA kernels use (closed-form rand() call index of every edge midpoint, levels breadth-first)."""
import json
import os

import numpy as np
import pytest

from tests import oracle_lib

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "terrain2d_golden.json")))


@pytest.mark.parametrize("depth,state", [(1, 1), (2, 12345), (3, 5), (4, 12345), (5, 31337), (6, 1), (7, 12345), (7, 0xFFFFFFFF), (8, 777), (9, 2)])
def test_rock_restatement_equals_reference(oracle, ref, depth, state):
    """faces {height, shade}, Cell attributes, rand() call count and generator state afterwards, bit for bit."""
    fr, cr, nr, sr = ref.rock(state, depth, 0.3)
    fo, co, no, so = oracle.rock(state, depth, 0.3)
    assert (no, so) == (nr, sr)
    assert no == oracle.L.rno_rock_draw_count(depth) == 4 * (1 << depth) ** 2 - 4
    assert np.array_equal(fr.view(np.uint32), fo.view(np.uint32))
    assert cr.tobytes() == co.tobytes()


def test_rock_amplitude_argument(oracle, ref):
    fr, cr, _, _ = ref.rock(42, 6, 0.45)
    fo, co, _, _ = oracle.rock(42, 6, 0.45)
    assert np.array_equal(fr.view(np.uint32), fo.view(np.uint32)) and cr.tobytes() == co.tobytes()


@pytest.mark.parametrize("entry", GOLD.get("rocks", []), ids=lambda e: f"d{e['depth']}-s{e['state']}")
def test_rock_golden(oracle, entry):
    faces, cells, draws, state_after = oracle.rock(entry["state"], entry["depth"], entry["amplitude"])
    assert draws == entry["rand_calls"] and state_after == entry["state_after"]
    assert oracle.fnv64(faces) == entry["fnv64_faces"]
    assert oracle.fnv64(cells.view(np.uint8)) == entry["fnv64_cells"]
    if "faces_bits" in entry:
        assert faces.view(np.uint32).reshape(-1).tolist() == entry["faces_bits"]
        assert cells.view(np.uint32).reshape(-1).tolist() == entry["cells_bits"]


@pytest.mark.parametrize("depth", [1, 2, 3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("state", [12345, 1, 0xDEADBEEF])
def test_rock_indexed_equals_sequential(oracle, depth, state):
    seq, _, draws, state_after = oracle.rock(state, depth, 0.3, want_cells=False)
    idx, n, st = oracle.rock_indexed(state, depth, 0.3)
    assert n == draws and st == state_after
    assert np.array_equal(seq.view(np.uint32), idx.view(np.uint32))


def test_rock_shared_edges_agree(oracle):
    """Every cube edge is seen by two faces through Rock::element; both views hold the same samples, and the eight
    cube corners are never displaced."""
    faces, _, _, _ = oracle.rock(12345, 5, 0.3, want_cells=False)
    W = 32
    # face 0: X = W, (Y, Z) = (x, y); face 2: Y = W, (X, Z) = (x, y): shared edge X = W, Y = W
    assert np.array_equal(faces[0, W, :, :].view(np.uint32), faces[2, W, :, :].view(np.uint32))
    for f in range(6):
        for x, y in ((0, 0), (0, W), (W, 0), (W, W)):
            assert faces[f, x, y, 0] == 0.0 and faces[f, x, y, 1] == 1.0


# ---- Cloud (Cloud.cpp), the two-face variant ----
@pytest.mark.parametrize("depth,state", [(2, 1), (3, 12345), (4, 5), (5, 12345), (6, 31337), (7, 12345), (8, 777), (9, 2)])
def test_cloud_restatement_equals_reference(oracle, ref, depth, state):
    fr, cr, nr, sr = ref.cloud(state, depth, 0.3)
    fo, co, no, so = oracle.cloud(state, depth, 0.3)
    assert (no, so) == (nr, sr)
    assert np.array_equal(fr.view(np.uint32), fo.view(np.uint32))
    assert cr.tobytes() == co.tobytes()


@pytest.mark.parametrize("entry", GOLD.get("clouds", []), ids=lambda e: f"d{e['depth']}-s{e['state']}")
def test_cloud_golden(oracle, entry):
    faces, cells, draws, state_after = oracle.cloud(entry["state"], entry["depth"], entry["amplitude"])
    assert draws == entry["rand_calls"] and state_after == entry["state_after"]
    assert oracle.fnv64(faces) == entry["fnv64_faces"]
    assert oracle.fnv64(cells.view(np.uint8)) == entry["fnv64_cells"]
    if "faces_bits" in entry:
        assert faces.view(np.uint32).reshape(-1).tolist() == entry["faces_bits"]
        assert cells.view(np.uint32).reshape(-1).tolist() == entry["cells_bits"]


def test_cloud_presets_and_shared_border(oracle):
    faces, _, draws, _ = oracle.cloud(12345, 5, 0.3, want_cells=False)
    W = 32
    for f, v in ((0, -0.5), (1, 0.5)):
        for x, y in ((W // 2, W // 2), (W // 4, W // 4), (3 * W // 4, W // 4), (W // 4, 3 * W // 4), (3 * W // 4, 3 * W // 4)):
            assert faces[f, x, y, 0] == v and faces[f, x, y, 1] == 1.0
    for sl in ((0, slice(None)), (W, slice(None)), (slice(None), 0), (slice(None), W)):
        assert np.array_equal(faces[0][sl].view(np.uint32), faces[1][sl].view(np.uint32))
    E = (W + 1) ** 2 - 4 - (4 ** 5 - 1) // 3
    assert draws == 2 * E - 4 * (W - 1)
