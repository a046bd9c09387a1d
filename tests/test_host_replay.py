"""CPU: the fused tile algorithm (rednoise_b200/csrc/rn_tile.cuh + rn_plan.h -- the very code the
CUDA stage kernels are compiled from) replayed thread by thread on the host
(tests/host_replay/rn_replay.cpp) against the oracle: tiling, halos, windows, boundary rules, RNG
indexing and shade ordering are validated bit for bit without a GPU.  The replay poisons the
tile's shared memory and the intermediate grids with NaN first, so a sample computed from
anything outside its dependency cone would show up as a NaN / bit mismatch."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_replay", "rn_replay.cpp")
SO = os.path.join(HERE, "host_replay", "librn_replay.so")
CSRC = os.path.join(os.path.dirname(HERE), "rednoise_b200", "csrc")


class Win(C.Structure):
    _fields_ = [("i0", C.c_int32), ("j0", C.c_int32), ("ni", C.c_int32), ("nj", C.c_int32)]


@pytest.fixture(scope="module")
def replay():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("rn_tile.cuh", "rn_stream.cuh", "rn_march.cuh", "rn_plan.h", "rn_math.cuh", "rn_volume.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", SO, SRC], check=True)
    L = C.CDLL(SO)
    L.replay_terrain2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    return L


def run(L, p, win=None, pitch=0, poison=1):
    s = (1 << p.depth) + 1
    ni, nj = (win.ni, win.nj) if win is not None else (s, s)
    out = np.full((ni, max(pitch, nj), 4), -7.0, np.float32)
    st = L.replay_terrain2d(C.byref(p), C.byref(win) if win is not None else None, out.ctypes.data, max(pitch, nj), poison)
    assert st == 0
    return out


@pytest.mark.parametrize("depth,seed", [(3, 12345), (4, 1), (5, 5), (6, 12345), (7, 70), (8, 12345), (8, 777), (9, 9), (10, 12345), (11, 3)])
@pytest.mark.parametrize("volcano", [0, 1])
def test_replay_equals_oracle_bitwise(replay, oracle, depth, seed, volcano):
    p = oracle_lib.params(seed=seed, depth=depth, apply_volcano=volcano)
    got = run(replay, p)
    want = oracle.terrain2d(p)
    assert not np.isnan(got).any()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("depth,seed", [(8, 5), (9, 12345), (10, 77)])
def test_replay_edge_instantiation_equals_interior(replay, oracle, depth, seed):
    """Every tile through the EDGE (existence-tested) instantiation == the mixed interior/edge run == the oracle."""
    p = oracle_lib.params(seed=seed, depth=depth, apply_volcano=1)
    a = run(replay, p, poison=3)
    assert np.array_equal(a.view(np.uint32), oracle.terrain2d(p).view(np.uint32))


@pytest.mark.parametrize("depth,seed", [(6, 3), (9, 12345), (11, 5)])
def test_replay_thread_order_does_not_matter(replay, oracle, depth, seed):
    """Threads of a phase visited in descending order give the same bits: no phase reads what another thread of
    the same phase writes (which would be a race between barriers on the GPU)."""
    p = oracle_lib.params(seed=seed, depth=depth, apply_volcano=1)
    a = run(replay, p, poison=5)
    assert np.array_equal(a.view(np.uint32), oracle.terrain2d(p).view(np.uint32))


def test_replay_depth12_digest(replay, oracle):
    """BASELINE config 2 against the reference digest (tests/golden)."""
    import json
    gold = json.load(open(os.path.join(HERE, "golden", "terrain2d_golden.json")))
    e = [f for f in gold["fields"] if f["depth"] == 12][0]
    got = run(replay, oracle_lib.params(seed=e["seed"], depth=12, apply_volcano=1))
    assert oracle.fnv64(got) == e["fnv64_post_volcano"]


def test_replay_nondefault_params(replay, oracle):
    kw = dict(persistence=0.6, sideways_scale=0.3, radiosity_strength=0.7, feature_scale=50.0, amplitude_variation=0.25,
              amplitude=0.35, bulgy_rocks=0)
    for depth in (5, 8):
        p = oracle_lib.params(seed=99, depth=depth, apply_volcano=1, **kw)
        assert np.array_equal(run(replay, p).view(np.uint32), oracle.terrain2d(p).view(np.uint32))


@pytest.mark.parametrize("kw", [
    dict(amplitude=5.0),                 # |amplitude| > 4: Law::flat_ok is off, the general tail everywhere
    dict(amplitude=0.02, size=80.0),     # a smooth field: the flat tail at every streamed level
    dict(amplitude=1.5, persistence=0.8, amplitude_variation=-0.9, sideways_scale=-0.5),  # rough: both tails
])
def test_replay_march_tails_with_nondefault_params(replay, oracle, kw):
    """rn_march.cuh pair_finish<true/false>: the flat tail (only z moves) and the general one reproduce the oracle for parameters
    that make every normal of the streamed levels degenerate, none of them, or that forbid the flat tail."""
    for depth, seed in ((10, 31), (11, 4)):
        p = oracle_lib.params(seed=seed, depth=depth, apply_volcano=1, **kw)
        assert np.array_equal(run(replay, p).view(np.uint32), oracle.terrain2d(p).view(np.uint32)), (kw, depth)


def test_replay_windows_are_seam_exact(replay, oracle):
    depth = 9
    S = (1 << depth) + 1
    p = oracle_lib.params(seed=4711, depth=depth, apply_volcano=1)
    full = oracle.terrain2d(p)
    rng = np.random.default_rng(5)
    wins = [(0, 0, 1, 1), (S - 1, S - 1, 1, 1), (0, 0, S, 1), (0, S - 1, S, 1), (100, 37, 64, 64), (255, 255, 3, 3), (0, 0, 129, S),
            (384, 0, S - 384, S), (64, 64, 64, 64), (63, 63, 2, 2), (512, 0, 1, S)]
    for _ in range(10):
        i0, j0 = (int(x) for x in rng.integers(0, S - 1, 2))
        wins.append((i0, j0, int(rng.integers(1, S - i0 + 1)), int(rng.integers(1, S - j0 + 1))))
    for (i0, j0, ni, nj) in wins:
        got = run(replay, p, Win(i0, j0, ni, nj), pitch=nj + 3)
        assert np.array_equal(got[:, :nj].view(np.uint32), full[i0:i0 + ni, j0:j0 + nj].view(np.uint32)), (i0, j0, ni, nj)
        assert (got[:, nj:] == -7.0).all()


def test_replay_streamed_levels_windows(replay, oracle):
    """Depth 10 uses the streaming kernel (rn_stream.cuh) for its last two levels: windows stay seam-exact."""
    depth = 10
    S = (1 << depth) + 1
    p = oracle_lib.params(seed=4711, depth=depth, apply_volcano=1)
    full = oracle.terrain2d(p)
    rng = np.random.default_rng(7)
    wins = [(100, 37, 300, 500), (513, 0, 512, S), (0, 0, 129, S), (64, 64, 64, 64), (63, 63, 130, 130), (512, 512, 1, 1), (200, 200, 700, 3)]
    for _ in range(5):
        i0, j0 = (int(x) for x in rng.integers(0, S - 1, 2))
        wins.append((i0, j0, int(rng.integers(1, S - i0 + 1)), int(rng.integers(1, S - j0 + 1))))
    for (i0, j0, ni, nj) in wins:
        got = run(replay, p, Win(i0, j0, ni, nj), pitch=nj + 3)
        assert np.array_equal(got[:, :nj].view(np.uint32), full[i0:i0 + ni, j0:j0 + nj].view(np.uint32)), (i0, j0, ni, nj)
        assert (got[:, nj:] == -7.0).all()


def test_stage_plans(replay):
    d, l = (C.c_int * 8)(), (C.c_int * 8)()
    for depth in range(3, 15):
        n = replay.replay_plan(depth, d, l)
        assert d[0] == 2 and sum(l[k] for k in range(n)) == depth - 2
        assert all(1 <= l[k] <= 4 for k in range(n)) and all(d[k + 1] == d[k] + l[k] for k in range(n - 1))
    for L in (1, 2, 3, 4):
        assert replay.replay_smem_bytes(L) <= 227 * 1024


def test_replay_march_depth12_windows_of_every_strip_layout(replay, oracle):
    """The warp march (rn_march.cuh) cuts a window into strips of 60 cell columns: one strip, strips that touch no border, only
    the left / right / top / bottom border, single rows and columns, odd pitches (sector-aligned and shifted rows) -- all seam-exact."""
    depth, seed = 12, 2718
    S = (1 << depth) + 1
    p = oracle_lib.params(seed=seed, depth=depth, apply_volcano=1)
    full = oracle.terrain2d(p)
    wins = [(1000, 1000, 300, 200), (1000, 1000, 64, 1500), (0, 1500, 500, 700), (S - 300, 2000, 300, 900), (1500, 0, 700, 260),
            (1500, S - 260, 700, 260), (0, 0, 257, 257), (S - 129, S - 129, 129, 129), (2047, 2047, 3, 3), (2048, 0, 1, S), (0, 2048, S, 1),
            (100, 3, 2200, 4000)]
    for k, (i0, j0, ni, nj) in enumerate(wins):
        got = run(replay, p, Win(i0, j0, ni, nj), pitch=nj + (k % 3), poison=1 if k % 2 else 5)
        assert np.array_equal(got[:, :nj].view(np.uint32), full[i0:i0 + ni, j0:j0 + nj].view(np.uint32)), (i0, j0, ni, nj)
        assert (got[:, nj:] == -7.0).all()


@pytest.mark.parametrize("depth,seed", [(10, 31), (11, 2024)])
def test_replay_block_streaming_kernel_agrees(replay, oracle, depth, seed):
    """poison bit 3: the streamed levels through the round-1 block kernel (rn_stream.cuh, algorithm 3) -- kept as a cross-check."""
    p = oracle_lib.params(seed=seed, depth=depth, apply_volcano=1)
    want = oracle.terrain2d(p)
    assert np.array_equal(run(replay, p, poison=9).view(np.uint32), want.view(np.uint32))
    assert np.array_equal(run(replay, p, poison=1).view(np.uint32), want.view(np.uint32))


def test_batch_chunk_rule(replay):
    """rn_plan.h march_batch_chunk: a chunk holds a field count for which every field's (strip, segment) units fill the warp slots and
    finish together -- never more than the memory bound allows, at least 4 fields when it allows that many, all of them for shallow fields."""
    f = replay.replay_batch_chunk
    assert f(10, 229, 148) == 27          # 7 interior strips x 8 segments + 2 border strips x 10 = 76 units x 27 fields = 2052 of 2072 slots
    for depth in (10, 11, 12, 13):
        for cmax in (1, 2, 3, 4, 7, 12, 50, 229, 1000):
            c = f(depth, cmax, 148)
            assert 1 <= c <= cmax and (c >= 4 or cmax < 4), (depth, cmax, c)
    assert f(9, 500, 148) == 500 and f(5, 32768, 148) == 32768   # tile stages only: the memory bound decides
    assert f(10, 229, 132) <= 229


def test_replay_march_is_the_default_for_deep_fields(replay):
    k = (C.c_int * 16)()
    for depth in (10, 12, 14):
        n = replay.replay_chain_kinds(depth, 1, k)
        assert list(k[:n])[-2:] == [2, 2]
        n = replay.replay_chain_kinds(depth, 0, k)
        assert list(k[:n])[-2:] == [1, 1]
    n = replay.replay_chain_kinds(9, 1, k)
    assert set(k[:n]) == {0}


def test_replay_march_deep_slab(replay, oracle):
    """A thin row slab and a column window of a depth-13 field (8193^2): strips far from every border, nseg > 1."""
    depth, seed = 13, 5
    p = oracle_lib.params(seed=seed, depth=depth, apply_volcano=1)
    full = oracle.terrain2d(p)
    for (i0, j0, ni, nj) in [(4000, 0, 200, 8193), (0, 4090, 8193, 70)]:
        got = run(replay, p, Win(i0, j0, ni, nj))
        assert np.array_equal(got.view(np.uint32), full[i0:i0 + ni, j0:j0 + nj].view(np.uint32)), (i0, j0, ni, nj)


def test_random_division_shortcut_is_exact(replay, oracle):
    """rn_math.cuh evaluates random()'s (2*r)/32767 as q = t*rc; q += fma(-q, 32767, t)*rc: the correctly rounded quotient, i.e. the
    oracle's (and the reference's) division, for EVERY one of the 32768 draws."""
    replay.replay_random_from_draw.restype = C.c_float
    replay.replay_random_from_draw.argtypes = [C.c_uint]
    for r in range(32768):
        a, b = replay.replay_random_from_draw(r), oracle.L.rno_random_from_draw(r)
        assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32), r
