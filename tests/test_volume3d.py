"""3D density-volume EXTENSION (BASELINE.json configs 512^3 / 2048^3).  The reference has no volume generator
(SURVEY.md section 0.2): parity is UNPINNED BY THE REFERENCE; the CPU specification oracle/rn_oracle3d.c is the
definition (round 2 changed its lattice hash, see that file's header; the digests below changed with it).
CPU tests: the specification's own invariants and committed digests.  GPU tests: the CUDA path reproduces the
specification bit for bit at every depth up to 513^3 (BASELINE configs[2]) and on thin slabs of the 2049^3 volume
(configs[3]); slabs are independent and slab-sized."""
import ctypes as C

import numpy as np
import pytest


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("depth", [2, 3, 4, 6])
def test_spec_is_complete_and_deterministic(oracle, depth):
    a = oracle.volume3d(12345, depth, 0.3)
    assert not np.isnan(a).any()
    assert np.array_equal(bits(a), bits(oracle.volume3d(12345, depth, 0.3)))
    assert not np.array_equal(bits(a), bits(oracle.volume3d(12346, depth, 0.3)))
    W = 1 << depth
    assert (a[::W // 2, ::W // 2, ::W // 2] == 0.0).all()          # the 27 undisplaced points
    assert abs(float(a.mean())) < 0.05 and 0.01 < float(a.std()) < 0.2


def test_spec_draws_are_15_bit_and_spread(oracle):
    """The integer lattice hash: 15-bit draws, uniform, and distinct along every axis -- also at the far corner of a 2049^3
    lattice, whose linear index no longer fits 32 bits (the hash takes the three coordinates, not the index)."""
    f = oracle.L.rno_volume_draw
    f.restype = C.c_uint32
    f.argtypes = [C.c_uint32] * 4
    rng = np.random.default_rng(3)
    pts = rng.integers(0, 2049, (20000, 3))
    d = np.array([f(7, int(i), int(j), int(k)) for i, j, k in pts])
    assert d.min() >= 0 and d.max() <= 0x7FFF and abs(d.mean() - 16383.5) < 300
    hist = np.bincount(d >> 11, minlength=16)
    assert hist.min() > 1000 and hist.max() < 1500
    row = [f(7, 2048, 2048, k) for k in range(2049)]
    assert len(set(row)) > 1900                                      # no structure along a row
    assert f(7, 1, 2, 3) != f(7, 3, 2, 1) and f(7, 1, 2, 3) != f(8, 1, 2, 3)

    def ref(seed, i, j, k):  # the definition, restated
        m = 0xFFFFFFFF
        h = (seed ^ (i * 0x9E3779B1) ^ (j * 0x85EBCA77) ^ (k * 0xC2B2AE3D)) & m
        h ^= h >> 16
        h = (h * 0x7FEB352D) & m
        h ^= h >> 15
        h = (h * 0x846CA68B) & m
        h ^= h >> 16
        return h >> 17
    for i, j, k in pts[:500]:
        assert f(12345, int(i), int(j), int(k)) == ref(12345, int(i), int(j), int(k))


@pytest.mark.parametrize("depth", [4, 6])
def test_spec_slabs_equal_the_full_volume(oracle, depth):
    S = (1 << depth) + 1
    full = oracle.volume3d(99, depth, 0.3)
    rng = np.random.default_rng(depth)
    cases = [(0, 1), (S - 1, 1), (0, S // 8 + 1), (S // 2, 3)] + [(int(a), int(rng.integers(1, S - a + 1))) for a in rng.integers(0, S, 5)]
    for i0, ni in cases:
        s = oracle.volume3d(99, depth, 0.3, i0, ni)
        assert s.shape == (ni, S, S)
        assert np.array_equal(bits(s), bits(full[i0:i0 + ni])), (i0, ni)


def test_spec_digests(oracle):
    got = {d: oracle.fnv64(oracle.volume3d(12345, d, 0.3)) for d in (3, 5, 6)}
    assert got == {3: "9343b089105a0c29", 5: "8485afe0ee91e0c0", 6: "18cf278bdefe007c"}


@pytest.fixture(scope="module")
def vreplay():
    """The 3D kernels' own tile routine (rednoise_b200/csrc/rn_volume.cuh) replayed on the host, one "thread" per tile, with the level
    grids and every tile's shared-memory image poisoned with NaN (tests/host_replay/rn_replay.cpp: replay_volume3d)."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    src = os.path.join(here, "host_replay", "rn_replay.cpp")
    so = os.path.join(here, "host_replay", "librn_replay.so")
    csrc = os.path.join(os.path.dirname(here), "rednoise_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("rn_tile.cuh", "rn_stream.cuh", "rn_march.cuh", "rn_plan.h", "rn_math.cuh", "rn_volume.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", so, src], check=True)
    L = C.CDLL(so)
    L.replay_volume3d.argtypes = [C.c_uint32, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]

    def run(seed, depth, i0=0, ni=None):
        S = (1 << depth) + 1
        ni = S - i0 if ni is None else ni
        v = np.full((ni, S, S), np.nan, np.float32)
        assert L.replay_volume3d(seed & 0xFFFFFFFF, depth, 0.3, i0, ni, v.ctypes.data, 1) == 0
        return v
    return run


@pytest.mark.parametrize("depth,seed", [(2, 1), (3, 12345), (4, 7), (5, 12345), (6, 99), (7, 2024), (8, 31337)])
def test_replay_volume_equals_spec_bitwise(vreplay, oracle, depth, seed):
    got = vreplay(seed, depth)
    assert not np.isnan(got).any()
    assert np.array_equal(bits(got), bits(oracle.volume3d(seed, depth, 0.3)))


def test_replay_volume_slabs(vreplay, oracle):
    depth = 7
    S = (1 << depth) + 1
    full = oracle.volume3d(5, depth, 0.3)
    for i0, ni in ((0, 1), (S - 1, 1), (0, 17), (16, 1), (15, 2), (31, 3), (64, 65), (100, 29), (17, 100)):
        assert np.array_equal(bits(vreplay(5, depth, i0, ni)), bits(full[i0:i0 + ni])), (i0, ni)


@pytest.mark.parametrize("i0,ni", [(0, 2), (1023, 3), (2047, 2)])
def test_replay_thin_slabs_of_the_2049_cubed_volume(vreplay, oracle, i0, ni):
    assert np.array_equal(bits(vreplay(777, 11, i0, ni)), bits(oracle.volume3d(777, 11, 0.3, i0, ni)))


def gpu_slab(ctx, seed, depth, i0=0, ni=None, stream=0):
    import torch
    S = (1 << depth) + 1
    ni = S - i0 if ni is None else ni
    v = torch.full((ni, S, S), float("nan"), device="cuda", dtype=torch.float32)
    ctx.volume3d_ptr(seed, depth, 0.3, v.data_ptr(), i0=i0, ni=ni, stream=stream)
    torch.cuda.synchronize()
    return v


@pytest.mark.gpu
@pytest.mark.parametrize("depth,seed", [(2, 1), (3, 12345), (4, 7), (5, 12345), (6, 99), (7, 2024), (8, 31337)])
def test_gpu_volume_equals_spec_bitwise(ctx, rnlib, oracle, depth, seed):
    got = gpu_slab(ctx, seed, depth).cpu().numpy()
    want = oracle.volume3d(seed, depth, 0.3)
    assert not np.isnan(got).any()
    assert np.array_equal(bits(got), bits(want))


@pytest.mark.gpu
def test_gpu_volume_513_cubed_equals_spec(ctx, rnlib, oracle):
    """BASELINE configs[2] (512^3, 8 'octaves') at full size against the specification."""
    got = gpu_slab(ctx, 12345, 9).cpu().numpy()
    want = oracle.volume3d(12345, 9, 0.3)
    assert np.array_equal(bits(got), bits(want))


@pytest.mark.gpu
def test_gpu_volume_slabs_are_independent_and_slab_sized(ctx, rnlib, oracle):
    """What the 2048^3 multi-GPU workload relies on: a slab generated alone, into a buffer of its own size, == the same planes
    of the whole volume."""
    import torch
    depth = 8
    S = (1 << depth) + 1
    full = gpu_slab(ctx, 5, depth)
    for world in (2, 8, 5):
        rows = (S + world - 1) // world
        for r in range(world):
            i0 = r * rows
            ni = min(rows, S - i0)
            if ni <= 0:
                continue
            slab = gpu_slab(ctx, 5, depth, i0, ni, stream=torch.cuda.current_stream().cuda_stream)
            assert torch.equal(slab.view(torch.int32), full[i0:i0 + ni].view(torch.int32)), (world, r)
    for i0, ni in ((0, 1), (S - 1, 1), (100, 1), (127, 3), (31, 2)):
        slab = gpu_slab(ctx, 5, depth, i0, ni)
        assert torch.equal(slab.view(torch.int32), full[i0:i0 + ni].view(torch.int32)), (i0, ni)


@pytest.mark.gpu
@pytest.mark.parametrize("i0,ni", [(0, 3), (1023, 3), (2046, 3), (1500, 2)])
def test_gpu_thin_slabs_of_the_2049_cubed_volume(ctx, rnlib, oracle, i0, ni):
    """BASELINE configs[3] (2048^3, 10 'octaves'): thin slabs -- first planes, around the middle, last planes (the far corner's
    linear index exceeds 2^32) -- against the specification's same slabs; neither side ever holds the 34 GB volume."""
    got = gpu_slab(ctx, 777, 11, i0, ni).cpu().numpy()
    want = oracle.volume3d(777, 11, 0.3, i0, ni)
    assert np.array_equal(bits(got), bits(want))


@pytest.mark.gpu
def test_gpu_back_to_back_async_volume_calls_do_not_mix(ctx, rnlib, oracle):
    """Call pipelining (rn_abi.cu rn_volume3d): the coarse levels of call k+1 run on a side stream into the other scratch slot while
    call k's last level is still running.  Five asynchronous calls with different seeds, depths and slabs on one stream, no
    synchronisation in between, interleaved with a 2D field that shares the scratch slots; every result against the specification."""
    import torch
    from tests import oracle_lib
    st = torch.cuda.Stream()
    jobs = [(11, 7, 0, None), (12, 7, 0, None), (13, 6, 0, None), (14, 7, 40, 50), (15, 5, 0, None), (16, 7, 0, None)]
    outs = []
    field = torch.empty((257, 257, 4), device="cuda", dtype=torch.float32)
    with torch.cuda.stream(st):
        for n, (seed, depth, i0, ni) in enumerate(jobs):
            S = (1 << depth) + 1
            o = torch.empty((S - i0 if ni is None else ni, S, S), device="cuda", dtype=torch.float32)
            ctx.volume3d_ptr(seed, depth, 0.3, o.data_ptr(), i0=i0, ni=ni, stream=st.cuda_stream)
            outs.append(o)
            if n == 2:
                ctx.terrain2d_ptr(rnlib.default_params(seed=5, depth=8, apply_volcano=1), field.data_ptr(), dst_is_device=True, stream=st.cuda_stream)
    st.synchronize()
    for (seed, depth, i0, ni), o in zip(jobs, outs):
        S = (1 << depth) + 1
        want = oracle.volume3d(seed, depth, 0.3, i0, S - i0 if ni is None else ni)
        assert np.array_equal(bits(o.cpu().numpy()), bits(want)), (seed, depth, i0, ni)
    want2d = oracle.terrain2d(oracle_lib.params(seed=5, depth=8, apply_volcano=1))
    assert np.array_equal(bits(field.cpu().numpy()), bits(want2d))


@pytest.mark.gpu
def test_gpu_volume_host_destination(ctx, rnlib, oracle):
    """rn_volume3d_host: the chunked, pipelined host path (pinned and pageable destinations, a slab that needs several chunks)
    equals the specification."""
    import torch
    depth, seed = 8, 77
    S = (1 << depth) + 1
    want = oracle.volume3d(seed, depth, 0.3)
    for pinned in (True, False):
        host = torch.empty((S, S, S), dtype=torch.float32, pin_memory=pinned)
        ctx.volume3d_to_host_ptr(seed, depth, 0.3, host.data_ptr())
        assert np.array_equal(host.numpy().view(np.uint32), want.view(np.uint32))
    part = torch.empty((5, S, S), dtype=torch.float32)
    ctx.volume3d_to_host_ptr(seed, depth, 0.3, part.data_ptr(), i0=100, ni=5)
    assert np.array_equal(part.numpy().view(np.uint32), want[100:105].view(np.uint32))
    L = rnlib.lib()
    p = rnlib.VolumeParams(seed, depth, 0.3)
    assert L.rn_volume3d_host(ctx._h, C.byref(p), 0, S + 1, C.c_void_p(part.data_ptr())) == rnlib.RN_ERR_BAD_PARAM
    assert L.rn_volume3d_host(ctx._h, C.byref(p), 0, 1, None) == rnlib.RN_ERR_BAD_PARAM


@pytest.mark.gpu
def test_gpu_volume_bad_arguments(ctx, rnlib):
    import torch
    L = rnlib.lib()
    v = torch.empty((5, 5, 5), device="cuda", dtype=torch.float32)
    p = rnlib.VolumeParams(1, 2, 0.3)
    assert L.rn_volume3d(ctx._h, C.byref(p), 0, 5, C.c_void_p(v.data_ptr()), None) == rnlib.RN_OK
    assert L.rn_volume3d(ctx._h, C.byref(p), 0, 6, C.c_void_p(v.data_ptr()), None) == rnlib.RN_ERR_BAD_PARAM
    assert L.rn_volume3d(ctx._h, C.byref(p), -1, 2, C.c_void_p(v.data_ptr()), None) == rnlib.RN_ERR_BAD_PARAM
    assert L.rn_volume3d(ctx._h, C.byref(rnlib.VolumeParams(1, 12, 0.3)), 0, 1, C.c_void_p(v.data_ptr()), None) == rnlib.RN_ERR_BAD_PARAM
    assert L.rn_volume3d(ctx._h, C.byref(p), 0, 5, None, None) == rnlib.RN_ERR_BAD_PARAM
    assert L.rn_volume3d_slab_elems(9, 2) == 2 * 513 * 513
