"""rednoise_b200 -- B200-native (sm_100a) red-noise field synthesis behind a C ABI.

The product is rednoise_b200/librednoise_b200.so (CUDA kernels + include/rednoise.h);
`capi` is the ctypes mirror of that header used by tests and bench.py.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
