"""Timeline of every launch of the stage chain over a few back-to-back calls (first CTA start / last CTA end, %globaltimer).

Needs an instrumented build (compiled out of the product):
    RN_NVCC_EXTRA=-DRN_TIMELINE python -m rednoise_b200.build --force
    python profiles/micro/timeline.py
    python -m rednoise_b200.build --force          # back to the product build
This is how DESIGN.md section 9 established what limits the overlap of consecutive calls."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
from rednoise_b200 import capi

ctx = capi.Context(0)
p = capi.default_params(seed=12345, depth=12, apply_volcano=1)
S = capi.side(12)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
fld = torch.empty((S, S, 4), device='cuda')
lib = capi.lib()
if not hasattr(lib, "rn_debug_timeline"):
    sys.exit("library was built without -DRN_TIMELINE")
lib.rn_debug_timeline.restype = ctypes.c_int
for _ in range(5): ctx.terrain2d_ptr(p, fld.data_ptr(), dst_is_device=True, stream=s.cuda_stream)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 2048)()
lib.rn_debug_timeline(buf, 1)
N = 8
for _ in range(N): ctx.terrain2d_ptr(p, fld.data_ptr(), dst_is_device=True, stream=s.cuda_stream)
torch.cuda.synchronize()
n = lib.rn_debug_timeline(buf, 0)
a = np.frombuffer(buf, dtype=np.uint64)
st = a[:n].astype(np.int64); en = a[1024:1024 + n].astype(np.int64)
t0 = st.min()
per = n // N
print("launches", n, "per call", per)
for k in range(n):
    print("call %2d stage %2d  start %8.1f  end %8.1f  dur %6.1f us" % (k // per, k % per, (st[k] - t0) / 1e3, (en[k] - t0) / 1e3, (en[k] - st[k]) / 1e3))
