"""Per-unit trace of the output kernel (k_march: one unit = one warp; algorithm 3, k_stream: one CTA): which SM ran each CTA, when it started and ended (%globaltimer).

Needs an instrumented build of the library (the instrumentation is compiled out of the product):
    RN_NVCC_EXTRA=-DRN_STREAM_TRACE python -m rednoise_b200.build --force
    python profiles/micro/cta_trace.py
    python -m rednoise_b200.build --force          # back to the product build
This is how DESIGN.md section 5.1 found that the single wave's duration was set by the border-strip CTAs (97 vs 78 us) and
that the hardware spreads the CTAs evenly (6 or 7 per SM)."""
import collections
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
from rednoise_b200 import capi

ctx = capi.Context(0)
if len(sys.argv) > 1:
    ctx.set_algorithm(int(sys.argv[1]))
p = capi.default_params(seed=12345, depth=12, apply_volcano=1)
S = capi.side(12)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
fld = torch.empty((S, S, 4), device='cuda')
for _ in range(4): ctx.terrain2d_ptr(p, fld.data_ptr(), dst_is_device=True, stream=s.cuda_stream)
torch.cuda.synchronize()
lib = capi.lib()
if not hasattr(lib, "rn_debug_trace"):
    sys.exit("library was built without -DRN_STREAM_TRACE")
n = 4 * 8192
buf = (ctypes.c_ulonglong * n)()
lib.rn_debug_trace.restype = ctypes.c_int
assert lib.rn_debug_trace(buf, n) == 0
a = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4)
a = a[a[:, 2] > 0]
# the trace holds the last launch that wrote each slot: the output kernel (1745 units) overwrites the lower slots of the last-but-one
# level (1904 units), whose remaining units are reported separately
older = a[a[:, 1] <= a[:, 1].max() - 20000]
if len(older):
    t0o = older[:, 1].min()
    d_o = (older[:, 2] - older[:, 1]).astype(np.int64)
    print("previous launch (last-but-one level), %d units still in the trace: duration us p10 %.1f p50 %.1f p90 %.1f max %.1f, span %.1f us" %
          ((len(older),) + tuple(np.percentile(d_o, [10, 50, 90, 100]) / 1e3) + ((older[:, 2].max() - t0o) / 1e3,)))
a = a[a[:, 1] > a[:, 1].max() - 20000]   # the last launch only (its CTAs start within 20 us of the latest start)
t0 = a[:, 1].min(); st = (a[:, 1] - t0).astype(np.int64); en = (a[:, 2] - t0).astype(np.int64); dur = en - st
print("ctas", len(a), " kernel span %.1f us  start spread p99 %.1f us" % (en.max() / 1e3, np.percentile(st, 99) / 1e3))
print("cta duration us: min %.1f p10 %.1f p50 %.1f p90 %.1f max %.1f" % tuple(np.percentile(dur, [0, 10, 50, 90, 100]) / 1e3))
e = a[:, 3] > 0
print("border-strip ctas %d: p50 %.1f max %.1f us; others: p50 %.1f max %.1f us" % (e.sum(), np.median(dur[e]) / 1e3, dur[e].max() / 1e3, np.median(dur[~e]) / 1e3, dur[~e].max() / 1e3))
per = collections.Counter(a[:, 0].tolist())
print("ctas per SM histogram", sorted(collections.Counter(per.values()).items()), "SMs used", len(per))
last = collections.defaultdict(int)
for smid, e_ in zip(a[:, 0].tolist(), en.tolist()): last[smid] = max(last[smid], e_)
l = np.array(list(last.values())) / 1e3
print("SM finish us: min %.1f p50 %.1f max %.1f" % (l.min(), np.median(l), l.max()))
print("cta end us percentiles 1,10,50,90,99:", np.percentile(en, [1, 10, 50, 90, 99]) / 1e3)
cnt = np.array([per[k] for k in sorted(per)])
fin = np.array([last[k] for k in sorted(per)]) / 1e3
for c in sorted(set(cnt.tolist())):
    print("  SMs with %d units: %d, finish p50 %.1f max %.1f us" % (c, (cnt == c).sum(), np.median(fin[cnt == c]), fin[cnt == c].max()))
print("start us percentiles 1,50,99,100:", np.percentile(st, [1, 50, 99, 100]) / 1e3)
