"""Builds rednoise_b200/librednoise_b200.so (CUDA kernels + the C ABI) in-tree for sm_100a.

    python -m rednoise_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  -fmad=false is REQUIRED: the FP32 field is
bit-exact with the reference only if a*b+c is never contracted.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librednoise_b200.so")
SOURCES = ["rn_abi.cu", "rn_levelsync.cu", "rn_util.cu", "rn_fused.cu", "rn_cells.cu", "rn_volume.cu", "rn_rock.cu"]
HEADERS = ["rn_math.cuh", "rn_kernels.h", "rn_tile.cuh", "rn_stream.cuh", "rn_march.cuh", "rn_volume.cuh", "rn_volrow.cuh", "rn_plan.h", os.path.join("..", "..", "include", "rednoise.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-fmad=false",            # no FMA contraction: bit-exact FP32 vs the x86 oracle
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xcompiler", "-ffp-contract=off",
    "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if os.environ.get("RN_LIB"):
        return os.environ["RN_LIB"]
    if not force and not _stale():
        return OUT
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    extra = os.environ.get("RN_NVCC_EXTRA", "").split()  # experiments, e.g. -DRN_STREAM_THREADS=160
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + srcs
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building librednoise_b200.so")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(OUT)
