"""GPU (B200): the CUDA path, called through the C ABI, against the CPU oracle on the same
seeds (bit-exact: integer draws AND FP32 fields), against the committed reference digests,
and -- at BASELINE.json's full sizes -- through size-independent properties (window/seam
exactness, batch == single, algorithm 1 == algorithm 2)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests import oracle_lib

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "terrain2d_golden.json")))
ALGOS = [1, 2]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def assert_bitwise(got, want, what=""):
    g, w = bits(got), bits(want)
    if not np.array_equal(g, w):
        bad = np.argwhere(g != w)
        i, j, c = bad[0]
        raise AssertionError(f"{what}: {len(bad)} of {g.size} words differ; first at (i={i}, j={j}, comp={c}): "
                             f"got {got[i, j, c]!r} want {want[i, j, c]!r}; "
                             f"pos words bad={int((g[..., :3] != w[..., :3]).sum())} shade words bad={int((g[..., 3] != w[..., 3]).sum())}")


def test_lcg_draws_bit_exact(ctx, oracle):
    """The integer lattice hash: the device draw stream equals the MSVC LCG everywhere."""
    for seed, first, n in [(1, 0, 64), (12345, 0, 100003), (777, 16785384 - 50, 50), (0xFFFFFFFF, 2**32 - 5, 37), (5, 3, 1)]:
        got = ctx.lcg_draws_host(seed, first, n)
        s = oracle.L.rno_lcg_state_after(seed, first)
        want = np.empty(n, np.uint16)
        for k in range(n):
            s = oracle.L.rno_lcg_next_state(s)
            want[k] = (s >> 16) & 0x7FFF
        assert np.array_equal(got, want), (seed, first, n)
    assert ctx.lcg_draws_host(1, 0, 8).tolist() == [41, 18467, 6334, 26500, 19169, 15724, 11478, 29358]


def test_normal_range_math_is_ieee_exact(ctx):
    """sqrt_rn_normal / rcp_rn_normal (fused kernels) == __fsqrt_rn / __frcp_rn on every binary32 in 2^-100..2^100."""
    assert ctx.selftest_math() == [0, 0, 0]


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("depth,seed", [(3, 12345), (4, 1), (5, 5), (6, 12345), (7, 70), (8, 12345), (8, 777), (9, 9), (10, 12345), (11, 3)])
@pytest.mark.parametrize("volcano", [0, 1])
def test_field_equals_oracle_bitwise(ctx, rnlib, oracle, algo, depth, seed, volcano):
    ctx.set_algorithm(algo)
    p = rnlib.default_params(seed=seed, depth=depth, apply_volcano=volcano)
    got, vpos = ctx.terrain2d_host(p, want_volcano_pos=True)
    want, wpos = oracle.terrain2d(oracle_lib.params(seed=seed, depth=depth, apply_volcano=volcano), want_volcano_pos=True)
    assert_bitwise(got, want, f"algo {algo} depth {depth} seed {seed} volcano {volcano}")
    assert vpos == wpos


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("entry", GOLD["fields"], ids=lambda e: f"d{e['depth']}-s{e['seed']}")
def test_field_matches_reference_digests(ctx, rnlib, oracle, algo, entry):
    """Includes BASELINE config 1 (depth 8) and config 2 (depth 12, 4097^2)."""
    ctx.set_algorithm(algo)
    for volc, key in ((0, "fnv64_pre_volcano"), (1, "fnv64_post_volcano")):
        p = rnlib.default_params(seed=entry["seed"], depth=entry["depth"], apply_volcano=volc)
        got = ctx.terrain2d_host(p)
        assert oracle.fnv64(got) == entry[key], (algo, entry["depth"], entry["seed"], volc)


@pytest.mark.parametrize("algo", ALGOS)
def test_nondefault_params(ctx, rnlib, oracle, algo):
    ctx.set_algorithm(algo)
    kw = dict(persistence=0.6, sideways_scale=0.3, radiosity_strength=0.7, feature_scale=50.0, amplitude_variation=0.25,
              amplitude=0.35, bulgy_rocks=0)
    for depth in (5, 8):
        got = ctx.terrain2d_host(rnlib.default_params(seed=99, depth=depth, apply_volcano=1, **kw))
        want = oracle.terrain2d(oracle_lib.params(seed=99, depth=depth, apply_volcano=1, **kw))
        assert_bitwise(got, want, f"non-default params depth {depth}")


@pytest.mark.parametrize("kw", [
    dict(persistence=0.6, sideways_scale=0.3, radiosity_strength=0.7, feature_scale=50.0, amplitude_variation=0.25, amplitude=0.35, bulgy_rocks=0),
    dict(amplitude=5.0),                 # |amplitude| > 4: Law::flat_ok is off, every warp takes the general tail
    dict(amplitude=0.02, size=80.0),     # a smooth field: the flat tail at every streamed level
    dict(amplitude=1.5, persistence=0.8, amplitude_variation=-0.9, sideways_scale=-0.5),  # rough: flat and general tails mixed
])
def test_deep_fields_with_nondefault_params(ctx, rnlib, oracle, kw):
    """The streamed levels' kernel (k_march, depth >= 10) on both tails of its per-warp decision (rn_march.cuh pair_finish): fields
    whose normals are all degenerate at the last levels, fields where none are, and parameters for which the flat tail is not allowed."""
    ctx.set_algorithm(0)
    for depth, seed in ((10, 31), (11, 4)):
        got = ctx.terrain2d_host(rnlib.default_params(seed=seed, depth=depth, apply_volcano=1, **kw))
        want = oracle.terrain2d(oracle_lib.params(seed=seed, depth=depth, apply_volcano=1, **kw))
        assert_bitwise(got, want, f"{kw} depth {depth}")


def test_large_host_destinations_pageable_and_pinned(ctx, rnlib, oracle):
    """rn_abi.cu enqueue_window_to_host: windows of >= 2^22 samples reach host memory in row chunks -- a pinned destination straight
    from the device, a PAGEABLE one (numpy / malloc / `new TerrainElement[]`) through two pinned halves and host copy threads.  Padded
    pitch, a window whose rows are no multiple of the chunk, several chunks, and the volcano position read back from the host copy."""
    import torch
    ctx.set_algorithm(0)
    depth, seed = 12, 2718
    S = (1 << depth) + 1
    full, wpos = oracle.terrain2d(oracle_lib.params(seed=seed, depth=depth, apply_volcano=1), want_volcano_pos=True)
    p = rnlib.default_params(seed=seed, depth=depth, apply_volcano=1)
    for (i0, j0, ni, nj, pad) in ((0, 0, S, S, 0), (3, 5, S - 10, S - 7, 9), (1000, 0, 1301, S, 0)):
        for pinned in (False, True):
            host = torch.full((ni, nj + pad, 4), -3.0, dtype=torch.float32, pin_memory=pinned)
            whole = (i0, j0, ni, nj) == (0, 0, S, S)
            vpos = ctx.terrain2d_ptr(p, host.data_ptr(), window=None if whole else rnlib.Window(i0, j0, ni, nj), pitch_elems=nj + pad,
                                     dst_is_device=False, want_volcano_pos=True)
            got = host.numpy()
            assert_bitwise(got[:, :nj], full[i0:i0 + ni, j0:j0 + nj], f"window {(i0, j0, ni, nj)} pinned={pinned}")
            assert (got[:, nj:] == -3.0).all()
            assert vpos == wpos


@pytest.mark.parametrize("algo", ALGOS)
def test_windows_are_seam_exact(ctx, rnlib, oracle, algo):
    """'grid extent and origin': any sub-window equals the same indices of the full field."""
    ctx.set_algorithm(algo)
    depth = 9
    p = rnlib.default_params(seed=4711, depth=depth, apply_volcano=1)
    full = oracle.terrain2d(oracle_lib.params(seed=4711, depth=depth, apply_volcano=1))
    S = (1 << depth) + 1
    rng = np.random.default_rng(5)
    wins = [(0, 0, 1, 1), (S - 1, S - 1, 1, 1), (0, 0, S, 1), (0, S - 1, S, 1), (100, 37, 64, 64), (255, 255, 3, 3), (0, 0, 129, S),
            (384, 0, S - 384, S)]
    for _ in range(6):
        i0, j0 = (int(x) for x in rng.integers(0, S - 1, 2))
        wins.append((i0, j0, int(rng.integers(1, S - i0 + 1)), int(rng.integers(1, S - j0 + 1))))
    for (i0, j0, ni, nj) in wins:
        got = ctx.terrain2d_host(p, window=rnlib.Window(i0, j0, ni, nj))
        assert_bitwise(got, full[i0:i0 + ni, j0:j0 + nj], f"window {(i0, j0, ni, nj)}")


def test_streamed_levels_windows_and_pitch(ctx, rnlib, oracle):
    """Depth 10/11 fields take the streaming kernel (k_stream) for their last two levels: windows of any shape,
    padded pitches and device destinations stay seam-exact (EDGE and lean strips, partial segments)."""
    import torch
    ctx.set_algorithm(0)
    for depth, seed in ((10, 4711), (11, 99)):
        S = (1 << depth) + 1
        p = rnlib.default_params(seed=seed, depth=depth, apply_volcano=1)
        full = oracle.terrain2d(oracle_lib.params(seed=seed, depth=depth, apply_volcano=1))
        rng = np.random.default_rng(depth)
        wins = [(0, 0, S, S), (100, 37, 300, 500), (S // 2, 0, S - S // 2, S), (0, 0, 129, S), (64, 64, 64, 64), (63, 63, 130, 130),
                (S // 2, S // 2, 1, 1), (200, 200, 700, 3), (S - 1, 0, 1, S), (0, S - 1, S, 1)]
        for _ in range(4):
            i0, j0 = (int(x) for x in rng.integers(0, S - 1, 2))
            wins.append((i0, j0, int(rng.integers(1, S - i0 + 1)), int(rng.integers(1, S - j0 + 1))))
        for (i0, j0, ni, nj) in wins:
            pitch = nj + 5
            d = torch.full((ni, pitch, 4), -1.0, device="cuda", dtype=torch.float32)
            ctx.terrain2d_ptr(p, d.data_ptr(), window=rnlib.Window(i0, j0, ni, nj), pitch_elems=pitch, dst_is_device=True,
                              stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            h = d.cpu().numpy()
            assert_bitwise(h[:, :nj], full[i0:i0 + ni, j0:j0 + nj], f"depth {depth} window {(i0, j0, ni, nj)}")
            assert (h[:, nj:] == -1.0).all()


def test_streamed_batch_and_slabs(ctx, rnlib, oracle):
    """BASELINE config 5 shape (depth-10 tiles, distinct seeds) through rn_terrain2d_batch, and a depth-12 field
    generated as 3 row slabs == the oracle's whole field (what the slab-sharded multi-GPU workload relies on)."""
    ctx.set_algorithm(0)
    seeds = [12345, 12346, 7, 0xFFFFFFFF]
    p = rnlib.default_params(depth=10, apply_volcano=1)
    got = ctx.terrain2d_batch_host(p, seeds)
    for t, s in enumerate(seeds):
        want = oracle.terrain2d(oracle_lib.params(seed=s, depth=10, apply_volcano=1))
        assert_bitwise(got[t], want, f"batch tile {t} seed {s}")
    from rednoise_b200 import sharding
    depth, world = 12, 3
    S = (1 << depth) + 1
    p12 = rnlib.default_params(seed=2024, depth=depth, apply_volcano=1)
    want = oracle.terrain2d(oracle_lib.params(seed=2024, depth=depth, apply_volcano=1))
    for r in range(world):
        i0, j0, ni, nj = sharding.slab_window(S, r, world)
        slab = ctx.terrain2d_host(p12, window=rnlib.Window(i0, j0, ni, nj))
        assert_bitwise(slab, want[i0:i0 + ni], f"slab {r} of {world}")


@pytest.mark.parametrize("depth,seed", [(3, 12345), (4, 9), (5, 7), (8, 12345), (10, 777), (11, 2024)])
def test_cell_quadtree_equals_oracle_bitwise(ctx, rnlib, oracle, depth, seed):
    """rn_terrain2d_cells (Terrain::generateRecursive on the GPU) vs the oracle: every byte of every rn_cell."""
    p = rnlib.default_params(seed=seed, depth=depth, apply_volcano=1)
    po = oracle_lib.params(seed=seed, depth=depth, apply_volcano=1)
    field = oracle.terrain2d(po)
    want = oracle.cells(po, field)
    got = ctx.cells_host(p, field)               # host field in, host cells out
    assert got.tobytes() == want.tobytes()
    got2 = ctx.cells_host(p)                     # field synthesized on the device first
    assert got2.tobytes() == want.tobytes()


def test_cell_quadtree_depth12_digest_and_device_path(ctx, rnlib, oracle):
    import torch
    gold = {e["depth"]: e for e in GOLD.get("cells", [])}
    p = rnlib.default_params(seed=12345, depth=12, apply_volcano=1)
    S = rnlib.side(12)
    n = int(rnlib.lib().rn_terrain2d_cell_count(12))
    fld = torch.empty((S, S, 4), device="cuda", dtype=torch.float32)
    cells = torch.empty((n, 8), device="cuda", dtype=torch.float32)
    stream = torch.cuda.current_stream().cuda_stream
    ctx.terrain2d_ptr(p, fld.data_ptr(), dst_is_device=True, stream=stream)
    ctx.cells_ptr(p, cells.data_ptr(), field_ptr=fld.data_ptr(), field_is_device=True, cells_is_device=True, stream=stream)
    torch.cuda.synchronize()
    po = oracle_lib.params(seed=12345, depth=12, apply_volcano=1)
    want = oracle.cells(po, oracle.terrain2d(po))
    assert cells.cpu().numpy().tobytes() == want.tobytes()
    # the depth-10 reference digest through the GPU path as well
    p10 = rnlib.default_params(seed=gold[10]["seed"], depth=10, apply_volcano=1)
    assert oracle.fnv64(ctx.cells_host(p10).view(np.uint8)) == gold[10]["fnv64"]


def test_many_seeds_and_depths(ctx, rnlib, oracle):
    """A sweep of seeds (the band / bulgy branches and the sqrt path of normalise are data dependent)."""
    ctx.set_algorithm(0)
    rng = np.random.default_rng(2026)
    cases = [(9, int(s)) for s in rng.integers(0, 2**32, 12)] + [(10, int(s)) for s in rng.integers(0, 2**32, 6)] + \
            [(11, int(s)) for s in rng.integers(0, 2**32, 2)]
    for depth, seed in cases:
        for volc in (0, 1):
            got = ctx.terrain2d_host(rnlib.default_params(seed=seed, depth=depth, apply_volcano=volc))
            want = oracle.terrain2d(oracle_lib.params(seed=seed, depth=depth, apply_volcano=volc))
            assert_bitwise(got, want, f"depth {depth} seed {seed} volcano {volc}")


def test_empty_batch_and_single_sample_windows(ctx, rnlib, oracle):
    import torch
    ctx.set_algorithm(0)
    p = rnlib.default_params(depth=10, apply_volcano=1)
    L = rnlib.lib()
    buf = np.zeros(4, np.float32)
    seeds = (C.c_uint32 * 1)(5)
    assert L.rn_terrain2d_batch(ctx._h, C.byref(p), seeds, 0, buf.ctypes.data, 0, 0, None) == rnlib.RN_OK   # ntiles == 0
    assert L.rn_terrain2d_batch(ctx._h, C.byref(p), seeds, -1, buf.ctypes.data, 0, 0, None) == rnlib.RN_ERR_BAD_PARAM
    full = oracle.terrain2d(oracle_lib.params(depth=10, apply_volcano=1))
    S = rnlib.side(10)
    for (i, j) in [(0, 0), (S - 1, S - 1), (S // 2, S // 2), (1, S - 2), (S - 1, 0)]:
        got = ctx.terrain2d_host(p, window=rnlib.Window(i, j, 1, 1))
        assert_bitwise(got, full[i:i + 1, j:j + 1], f"1x1 window at {(i, j)}")


def test_bad_arguments(ctx, rnlib):
    L = rnlib.lib()
    buf = np.empty((9 * 9, 4), np.float32)
    p = rnlib.default_params(depth=3)
    def call(pp, win=None, pitch=0):
        return L.rn_terrain2d(ctx._h, C.byref(pp), C.byref(win) if win else None, buf.ctypes.data, pitch, 0, None, None)
    assert call(p) == rnlib.RN_OK
    assert call(rnlib.default_params(depth=2)) == rnlib.RN_ERR_BAD_PARAM
    assert call(rnlib.default_params(depth=15)) == rnlib.RN_ERR_BAD_PARAM
    assert call(rnlib.default_params(depth=3, lacunarity=3)) == rnlib.RN_ERR_UNSUPPORTED
    assert call(rnlib.default_params(depth=3, persistence=-1.0)) == rnlib.RN_ERR_BAD_PARAM
    assert call(p, rnlib.Window(0, 0, 10, 9)) == rnlib.RN_ERR_BAD_PARAM
    assert call(p, rnlib.Window(-1, 0, 2, 2)) == rnlib.RN_ERR_BAD_PARAM
    assert call(p, rnlib.Window(0, 0, 0, 2)) == rnlib.RN_ERR_BAD_PARAM
    assert call(p, rnlib.Window(0, 0, 4, 4), pitch=3) == rnlib.RN_ERR_BAD_PARAM
    assert L.rn_terrain2d(ctx._h, C.byref(p), None, None, 0, 0, None, None) == rnlib.RN_ERR_BAD_PARAM
    assert L.rn_last_error_string() != b""


@pytest.mark.parametrize("algo", ALGOS)
def test_device_destination_and_pitch(ctx, rnlib, oracle, algo):
    import torch
    ctx.set_algorithm(algo)
    depth = 7
    S = (1 << depth) + 1
    want = oracle.terrain2d(oracle_lib.params(seed=8, depth=depth, apply_volcano=1))
    p = rnlib.default_params(seed=8, depth=depth, apply_volcano=1)
    # dense, in place
    d = torch.full((S, S, 4), float("nan"), device="cuda", dtype=torch.float32)
    ctx.terrain2d_ptr(p, d.data_ptr(), dst_is_device=True)
    assert_bitwise(d.cpu().numpy(), want, "device dense")
    # padded pitch, on torch's current stream (asynchronous form)
    pitch = S + 7
    d2 = torch.full((S, pitch, 4), -1.0, device="cuda", dtype=torch.float32)
    ctx.terrain2d_ptr(p, d2.data_ptr(), pitch_elems=pitch, dst_is_device=True, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    h = d2.cpu().numpy()
    assert_bitwise(h[:, :S], want, "device pitched")
    assert (h[:, S:] == -1.0).all()
    # window into a device buffer
    d3 = torch.zeros((20, 30, 4), device="cuda", dtype=torch.float32)
    ctx.terrain2d_ptr(p, d3.data_ptr(), window=rnlib.Window(5, 9, 20, 30), dst_is_device=True)
    assert_bitwise(d3.cpu().numpy(), want[5:25, 9:39], "device window")


@pytest.mark.parametrize("algo", ALGOS)
def test_batch_equals_singles(ctx, rnlib, oracle, algo):
    """BASELINE config 5 in miniature: tile k of a batch == a single run with seed k."""
    ctx.set_algorithm(algo)
    depth = 7
    seeds = [12345 + k for k in range(9)] + [0, 0xFFFFFFFF]
    p = rnlib.default_params(depth=depth, apply_volcano=1)
    got = ctx.terrain2d_batch_host(p, seeds)
    want = oracle.terrain2d_batch(oracle_lib.params(depth=depth, apply_volcano=1), seeds, threads=4)
    for t in range(len(seeds)):
        assert_bitwise(got[t], want[t], f"batch tile {t}")


def test_algorithms_agree_at_full_size(ctx, rnlib, oracle):
    """Depth 12 (4097^2), 13 and 14 (16385^2, the deepest field the ABI accepts): the streamed/fused path and the
    level-synchronous path are independent implementations; bitwise equality at sizes the oracle is slow for."""
    import torch
    for depth, seed in ((12, 2), (13, 12345), (14, 99)):
        S = (1 << depth) + 1
        p = rnlib.default_params(seed=seed, depth=depth, apply_volcano=1)
        a = torch.empty((S, S, 4), device="cuda", dtype=torch.float32)
        b = torch.empty((S, S, 4), device="cuda", dtype=torch.float32)
        ctx.set_algorithm(1)
        ctx.terrain2d_ptr(p, a.data_ptr(), dst_is_device=True)
        ctx.set_algorithm(2)
        ctx.terrain2d_ptr(p, b.data_ptr(), dst_is_device=True)
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), depth
        del a, b


def test_depth12_against_oracle(ctx, rnlib, oracle):
    """BASELINE config 2 against the oracle itself (about 1 s of CPU)."""
    ctx.set_algorithm(0)
    p = rnlib.default_params(seed=31337, depth=12, apply_volcano=1)
    got = ctx.terrain2d_host(p)
    want = oracle.terrain2d(oracle_lib.params(seed=31337, depth=12, apply_volcano=1))
    assert_bitwise(got, want, "depth 12 seed 31337")


def test_depth12_windows_of_every_strip_layout(ctx, rnlib, oracle):
    """The streamed output level cuts a window into column strips (border strips into shorter segments): one strip, strips that
    touch no border, only the left / right / top / bottom border, single rows and columns -- all seam-exact against the whole field."""
    ctx.set_algorithm(0)
    depth, seed = 12, 2718
    S = (1 << depth) + 1
    p = rnlib.default_params(seed=seed, depth=depth, apply_volcano=1)
    full = oracle.terrain2d(oracle_lib.params(seed=seed, depth=depth, apply_volcano=1))
    wins = [(1000, 1000, 300, 200), (1000, 1000, 64, 1500), (0, 1500, 500, 700), (S - 300, 2000, 300, 900), (1500, 0, 700, 260),
            (1500, S - 260, 700, 260), (0, 0, 257, 257), (S - 129, S - 129, 129, 129), (2047, 2047, 3, 3), (2048, 0, 1, S), (0, 2048, S, 1),
            (100, 3, 2200, 4000)]
    for (i0, j0, ni, nj) in wins:
        got = ctx.terrain2d_host(p, window=rnlib.Window(i0, j0, ni, nj))
        assert_bitwise(got, full[i0:i0 + ni, j0:j0 + nj], f"depth 12 window {(i0, j0, ni, nj)}")


def test_deep_batches_of_one_and_two(ctx, rnlib, oracle):
    """One deep field per call takes the small coarse chain (32x32 tiles, 128-thread CTAs), two or more the 64x64 tile stages:
    both through the batch entry point (seeds from device memory), device destination, against the oracle."""
    import torch
    ctx.set_algorithm(0)
    depth = 11
    S = (1 << depth) + 1
    p = rnlib.default_params(depth=depth, apply_volcano=1)
    for seeds in ([777], [31337, 4]):
        d = torch.empty((len(seeds), S, S, 4), device="cuda", dtype=torch.float32)
        ctx.terrain2d_batch_ptr(p, seeds, d.data_ptr(), dst_is_device=True)
        torch.cuda.synchronize()
        got = d.cpu().numpy()
        for t, seed in enumerate(seeds):
            want = oracle.terrain2d(oracle_lib.params(seed=seed, depth=depth, apply_volcano=1))
            assert_bitwise(got[t], want, f"batch of {len(seeds)}, tile {t}")
    # back-to-back single-field calls (call pipelining: the next call's coarse chain runs under this call's output kernel)
    outs = [torch.empty((S, S, 4), device="cuda", dtype=torch.float32) for _ in range(4)]
    s = torch.cuda.Stream()
    for k, o in enumerate(outs):
        ctx.terrain2d_ptr(rnlib.default_params(seed=100 + k, depth=depth, apply_volcano=1), o.data_ptr(), dst_is_device=True, stream=s.cuda_stream)
    torch.cuda.synchronize()
    for k, o in enumerate(outs):
        want = oracle.terrain2d(oracle_lib.params(seed=100 + k, depth=depth, apply_volcano=1))
        assert_bitwise(o.cpu().numpy(), want, f"pipelined call {k}")


def test_async_batches_with_pinned_seeds_do_not_mix(ctx, rnlib, oracle):
    """Three back-to-back asynchronous rn_terrain2d_batch calls, different seeds in PINNED host arrays, no synchronisation in
    between: the seeds upload of a call is truly asynchronous then, and the call's first stages run on the context's side
    streams -- they must wait for the upload and must not see the previous call's seeds (round-1 ADVICE, rn_abi.cu)."""
    import torch
    ctx.set_algorithm(0)
    for depth, ntiles in ((10, 3), (8, 5)):
        S = (1 << depth) + 1
        p = rnlib.default_params(depth=depth, apply_volcano=1)
        calls = 3
        seeds = [torch.empty((ntiles,), dtype=torch.int32).pin_memory() for _ in range(calls)]
        for k in range(calls):
            for t in range(ntiles):
                seeds[k][t] = 1000 * (k + 1) + 17 * t
        outs = [torch.empty((ntiles, S, S, 4), device="cuda", dtype=torch.float32) for _ in range(calls)]
        s = torch.cuda.Stream()
        # keep the stream busy first so that the uploads really queue behind device work
        busy = torch.empty((1 << 26,), device="cuda", dtype=torch.float32)
        with torch.cuda.stream(s):
            for _ in range(4):
                busy.add_(1.0)
        for k in range(calls):
            ctx.terrain2d_batch_ptr(p, ntiles, outs[k].data_ptr(), dst_is_device=True, stream=s.cuda_stream, seeds_ptr=seeds[k].data_ptr())
        torch.cuda.synchronize()
        for k in range(calls):
            got = outs[k].cpu().numpy()
            for t in range(ntiles):
                want = oracle.terrain2d(oracle_lib.params(seed=int(seeds[k][t]), depth=depth, apply_volcano=1))
                assert_bitwise(got[t], want, f"depth {depth} call {k} tile {t}")


def test_block_streaming_kernel_still_agrees(ctx, rnlib, oracle):
    """Algorithm 3 (the round-1 block streaming kernel for the last two levels) is kept as an independent implementation of what
    the warp march (algorithm 2, the default) computes: same bits at depth 10-13, and on a window."""
    import torch
    for depth, seed in ((10, 5), (11, 77), (12, 12345), (13, 9)):
        S = (1 << depth) + 1
        p = rnlib.default_params(seed=seed, depth=depth, apply_volcano=1)
        a = torch.empty((S, S, 4), device="cuda", dtype=torch.float32)
        b = torch.empty((S, S, 4), device="cuda", dtype=torch.float32)
        ctx.set_algorithm(3)
        ctx.terrain2d_ptr(p, a.data_ptr(), dst_is_device=True)
        ctx.set_algorithm(2)
        ctx.terrain2d_ptr(p, b.data_ptr(), dst_is_device=True)
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), depth
        del a, b
    ctx.set_algorithm(3)
    p = rnlib.default_params(seed=4711, depth=10, apply_volcano=1)
    full = oracle.terrain2d(oracle_lib.params(seed=4711, depth=10, apply_volcano=1))
    got = ctx.terrain2d_host(p, window=rnlib.Window(100, 37, 300, 500))
    assert_bitwise(got, full[100:400, 37:537], "algorithm 3 window")
    ctx.set_algorithm(0)


def test_march_even_pitch_and_odd_alignment(ctx, rnlib, oracle):
    """The march kernel stores whole 32-byte sectors; which rows are sector-aligned depends on the destination's base address
    and pitch parity.  Full fields into destinations with even / odd pitch and a base shifted by one element."""
    import torch
    ctx.set_algorithm(0)
    depth, seed = 11, 321
    S = (1 << depth) + 1
    p = rnlib.default_params(seed=seed, depth=depth, apply_volcano=1)
    want = oracle.terrain2d(oracle_lib.params(seed=seed, depth=depth, apply_volcano=1))
    for pitch, shift in ((S, 0), (S + 1, 0), (S, 1), (S + 1, 1), (S + 7, 3)):
        buf = torch.full((S * pitch + 8, 4), -3.0, device="cuda", dtype=torch.float32)
        ctx.terrain2d_ptr(p, buf.data_ptr() + 16 * shift, pitch_elems=pitch, dst_is_device=True)
        h = buf.cpu().numpy()[shift:shift + S * pitch].reshape(S, pitch, 4)
        assert_bitwise(h[:, :S], want, f"pitch {pitch} shift {shift}")
        assert (h[:, S:] == -3.0).all()
