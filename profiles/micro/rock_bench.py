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

# the app's rock loop (Core.cpp:69-79): 20 rocks of depth 2..7 continuing one rand() stream -- rn_rock_batch (ONE kernel launch for
# the batch) against 20 rn_rock calls, device destinations
import ctypes as C
L = capi.lib()
state, nrocks = 12345, 20
s = state
depths = []
for _ in range(nrocks):
    s = (s * 214013 + 2531011) & 0xFFFFFFFF
    e = np.float32(-5.0) + np.float32(5.0) * (np.float32((s >> 16) & 0x7FFF) / np.float32(32767.0))
    depths.append(int(np.float32(e + np.float32(7.0))))
    s = int(L.rn_lcg_state_after(s, int(L.rn_rock_draw_count(depths[-1]))))
fbuf = [torch.empty((6, (1 << d) + 1, (1 << d) + 1, 2), device='cuda') for d in depths]
cbuf = [torch.empty((6, int(L.rn_terrain2d_cell_count(d)), 8), device='cuda') for d in depths]
fp = (C.c_void_p * nrocks)(*[t.data_ptr() for t in fbuf])
cp = (C.c_void_p * nrocks)(*[t.data_ptr() for t in cbuf])
after = C.c_uint32(0)


def batch():
    st = L.rn_rock_batch(ctx._h, state, nrocks, None, 0.3, fp, cp, 1, None, None, None, None, C.byref(after))
    assert st == 0, st


def one_by_one():
    s = state
    for i, d in enumerate(depths):
        s = (s * 214013 + 2531011) & 0xFFFFFFFF
        s = ctx.rock_ptr(s, d, 0.3, fbuf[i].data_ptr(), cbuf[i].data_ptr())
    return s


for fn, name in ((batch, "rn_rock_batch (one launch)"), (one_by_one, "20 x rn_rock")):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print("20 rocks of depth %s: %-28s %8.1f us (median of 10, wall clock incl. the status read-back)" % (depths, name, sorted(ts)[5] * 1e6))
assert one_by_one() == after.value
