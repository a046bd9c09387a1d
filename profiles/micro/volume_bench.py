# 3D extension timing: python profiles/micro/volume_bench.py [depth]   (depth 9 = 512^3)
import sys, time, torch
sys.path.insert(0, '/root/repo')
from rednoise_b200 import capi
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 9
S = (1 << depth) + 1
ctx = capi.Context(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
v = torch.empty((S, S, S), device='cuda', dtype=torch.float32)
for _ in range(3): ctx.volume3d_ptr(12345, depth, 0.3, v.data_ptr(), stream=s.cuda_stream)
torch.cuda.synchronize()
n = 8
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
l0 = ctx.launches
ev[0].record()
for k in range(n):
    ctx.volume3d_ptr(12345, depth, 0.3, v.data_ptr(), stream=s.cuda_stream); ev[k + 1].record()
torch.cuda.synchronize()
ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(n))[n // 2]
print("volume %d^3 (%d levels): median %.3f ms, %.1f Gsamples/s, %.0f GB/s of finished samples (4 B each), %d launches per volume" %
      (S, depth - 1, ms, S**3 / ms / 1e6, S**3 * 4 / ms / 1e6, (ctx.launches - l0) // n))
