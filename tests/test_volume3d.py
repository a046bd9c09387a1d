"""3D density-volume EXTENSION (BASELINE.json configs 512^3 / 2048^3).  The reference has no volume generator
(SURVEY.md section 0.2): parity is UNPINNED BY THE REFERENCE; the CPU specification oracle/rn_oracle3d.c is the
definition.  CPU tests: the specification's own invariants and committed digests.  GPU tests: the CUDA path
reproduces the specification bit for bit, slabs are independent."""
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
    oracle.L.rno_volume_draw.restype = C.c_uint32
    oracle.L.rno_volume_draw.argtypes = [C.c_uint32, C.c_uint64]
    d = np.array([oracle.L.rno_volume_draw(7, k) for k in range(20000)])
    assert d.min() >= 0 and d.max() <= 0x7FFF and abs(d.mean() - 16383.5) < 300
    assert oracle.L.rno_volume_draw(7, 2**33 + 5) != oracle.L.rno_volume_draw(7, 5)   # indices beyond 2^32 are distinct draws


@pytest.mark.parametrize("depth", [4, 6])
def test_spec_slabs_equal_the_full_volume(oracle, depth):
    S = (1 << depth) + 1
    full = oracle.volume3d(99, depth, 0.3)
    rng = np.random.default_rng(depth)
    cases = [(0, 1), (S - 1, 1), (0, S // 8 + 1), (S // 2, 3)] + [(int(a), int(rng.integers(1, S - a + 1))) for a in rng.integers(0, S, 5)]
    for i0, ni in cases:
        s = oracle.volume3d(99, depth, 0.3, i0, ni)
        assert np.array_equal(bits(s[i0:i0 + ni]), bits(full[i0:i0 + ni])), (i0, ni)


def test_spec_digests(oracle):
    got = {d: oracle.fnv64(oracle.volume3d(12345, d, 0.3)) for d in (3, 5, 6)}
    assert got == {3: "79a85bd3e0d5df38", 5: "abebeb0f132b9a31", 6: "787b6f4aa137642c"}


@pytest.mark.gpu
@pytest.mark.parametrize("depth,seed", [(2, 1), (3, 12345), (4, 7), (5, 12345), (6, 99), (7, 2024), (8, 31337)])
def test_gpu_volume_equals_spec_bitwise(ctx, rnlib, oracle, depth, seed):
    import torch
    S = (1 << depth) + 1
    v = torch.full((S, S, S), float("nan"), device="cuda", dtype=torch.float32)
    ctx.volume3d_ptr(seed, depth, 0.3, v.data_ptr())
    want = oracle.volume3d(seed, depth, 0.3)
    got = v.cpu().numpy()
    assert not np.isnan(got).any()
    assert np.array_equal(bits(got), bits(want))


@pytest.mark.gpu
def test_gpu_volume_slabs_are_independent(ctx, rnlib, oracle):
    """What the 2048^3 multi-GPU workload relies on: a slab generated alone == the same planes of the whole volume."""
    import torch
    depth = 8
    S = (1 << depth) + 1
    full = torch.empty((S, S, S), device="cuda", dtype=torch.float32)
    ctx.volume3d_ptr(5, depth, 0.3, full.data_ptr())
    for world in (2, 8, 5):
        rows = (S + world - 1) // world
        for r in range(world):
            i0 = r * rows
            ni = min(rows, S - i0)
            if ni <= 0:
                continue
            slab = torch.full((S, S, S), float("nan"), device="cuda", dtype=torch.float32)
            ctx.volume3d_ptr(5, depth, 0.3, slab.data_ptr(), i0=i0, ni=ni, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert torch.equal(slab[i0:i0 + ni].view(torch.int32), full[i0:i0 + ni].view(torch.int32)), (world, r)


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
