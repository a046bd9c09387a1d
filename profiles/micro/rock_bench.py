"""rn_rock timing beside the CPU restatement of Rock.cpp (one host thread, like the reference):
    python profiles/micro/rock_bench.py
Device destination; the call synchronises (it returns the sentinel status), so wall clock around the call is the latency."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '/root/repo')
from rednoise_b200 import capi
from tests import oracle_lib

ctx = capi.Context(0)
o = oracle_lib.load_oracle()
for depth in (5, 7, 8, 10, 11):
    side = (1 << depth) + 1
    n = int(capi.lib().rn_terrain2d_cell_count(depth))
    faces = torch.empty((6, side, side, 2), device='cuda'); cells = torch.empty((6, n, 8), device='cuda')
    for _ in range(3): ctx.rock_ptr(12345, depth, 0.3, faces.data_ptr(), cells.data_ptr())
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); ctx.rock_ptr(12345, depth, 0.3, faces.data_ptr(), cells.data_ptr()); ts.append(time.perf_counter() - t0)
    g = sorted(ts)[len(ts) // 2]
    t0 = time.perf_counter(); fo, co, draws, so = o.rock(12345, depth, 0.3); c = time.perf_counter() - t0
    same = np.array_equal(fo.view(np.uint32), faces.cpu().numpy().view(np.uint32)) and co.tobytes() == cells.cpu().numpy().tobytes()
    samples = 6 * side * side
    print("rock depth %2d: %9d samples + %8d cells  GPU %8.1f us (%6.2f Gsamples/s)   CPU oracle %8.2f ms (%.4f Gsamples/s)   identical: %s"
          % (depth, samples, 6 * n, g * 1e6, samples / g / 1e9, c * 1e3, samples / c / 1e9, same))
