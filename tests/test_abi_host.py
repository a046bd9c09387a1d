"""CPU: the C-ABI library loads, exports every symbol include/rednoise.h declares, and its
host-only helpers agree with the oracle.  No compute call is made (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rednoise.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(rnlib):
    L = rnlib.lib()
    names = declared_symbols()
    assert sorted(names) == sorted(rnlib.SYMBOLS)
    for n in names:
        assert hasattr(L, n), n


def test_struct_layouts(rnlib):
    assert C.sizeof(rnlib.Params) == 48
    assert C.sizeof(rnlib.Window) == 16


def test_defaults_are_the_reference_literals(rnlib):
    p = rnlib.default_params()
    assert (p.depth, p.lacunarity, p.bulgy_rocks, p.apply_volcano) == (12, 2, 1, 1)
    assert (p.amplitude, p.size, p.persistence) == (np.float32(0.2), 500.0, 0.5)
    assert (p.sideways_scale, p.radiosity_strength, p.feature_scale, p.amplitude_variation) == (0.5, 0.5, 80.0, 0.5)


def test_host_helpers_match_oracle(rnlib, oracle):
    L = rnlib.lib()
    assert L.rn_version() == 100
    for d in range(3, 15):
        assert L.rn_terrain2d_side(d) == (1 << d) + 1
        assert L.rn_terrain2d_draw_count(d) == oracle.L.rno_draw_count(d)
    assert L.rn_terrain2d_side(2) == -1 and L.rn_terrain2d_side(15) == -1
    rng = np.random.default_rng(1)
    for _ in range(300):
        seed, k = int(rng.integers(0, 2**32)), int(rng.integers(0, 2**34))
        assert L.rn_lcg_state_after(seed, k) == oracle.L.rno_lcg_state_after(seed, k)
        assert L.rn_lcg_draw(seed, k) == oracle.L.rno_lcg_draw(seed, k)
    for d in (3, 4, 6):
        W = 1 << d
        for i in range(W + 1):
            for j in range(W + 1):
                assert L.rn_terrain2d_call_index(d, i, j) == oracle.L.rno_call_index(d, i, j)
    for _ in range(2000):
        d = int(rng.integers(3, 15))
        i, j = (int(x) for x in rng.integers(0, (1 << d) + 1, 2))
        assert L.rn_terrain2d_call_index(d, i, j) == oracle.L.rno_call_index(d, i, j)
    assert L.rn_terrain2d_call_index(5, -1, 0) == 2**64 - 1


def test_no_cpu_fallback(rnlib):
    """Without a device the compute entry points refuse to run; they never fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rnlib.RedNoiseError) as e:
        rnlib.Context(0)
    assert e.value.status == rnlib.RN_ERR_NO_DEVICE
    L = rnlib.lib()
    p = rnlib.default_params()
    buf = np.empty(16, np.float32)
    assert L.rn_terrain2d(None, C.byref(p), None, buf.ctypes.data, 0, 0, None, None) == rnlib.RN_ERR_BAD_PARAM
    assert b"ctx" in L.rn_last_error_string()


def test_product_never_touches_oracle():
    """The shipped package must not import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "rednoise_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".c")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                for token in ("oracle/", "rn_oracle", "librn_oracle", "libredref", "oracle_lib", "rno_"):
                    assert token not in src, (os.path.join(dirpath, f), token)
