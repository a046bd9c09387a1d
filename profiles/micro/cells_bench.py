import sys, torch, time
sys.path.insert(0, '/root/repo')
from rednoise_b200 import capi
ctx = capi.Context(0)
p = capi.default_params(seed=12345, depth=12, apply_volcano=1)
S = capi.side(12); n = int(capi.lib().rn_terrain2d_cell_count(12))
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
fld = torch.empty((S, S, 4), device='cuda'); cells = torch.empty((n, 8), device='cuda')
ctx.terrain2d_ptr(p, fld.data_ptr(), dst_is_device=True, stream=s.cuda_stream)
for _ in range(3): ctx.cells_ptr(p, cells.data_ptr(), field_ptr=fld.data_ptr(), stream=s.cuda_stream)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
ev[0].record()
for k in range(10):
    ctx.cells_ptr(p, cells.data_ptr(), field_ptr=fld.data_ptr(), stream=s.cuda_stream); ev[k+1].record()
torch.cuda.synchronize()
ms = [ev[k].elapsed_time(ev[k+1]) for k in range(10)]
print("k_cells depth 12: %d cells, median %.1f us, %.1f Gcells/s, algorithmic bytes %.0f MB -> %.0f GB/s" % (n, 1e3*sorted(ms)[5], n/sorted(ms)[5]/1e6, (S*S*16+n*32)/1e6, (S*S*16+n*32)/sorted(ms)[5]/1e6))
# CPU beside it: the oracle restatement of Terrain::generateRecursive (one host thread, like the reference)
import numpy as np
from tests import oracle_lib
o = oracle_lib.load_oracle()
po = oracle_lib.params(seed=12345, depth=12, apply_volcano=1)
f = o.terrain2d(po)
t0 = time.perf_counter(); c = o.cells(po, f); dt = time.perf_counter() - t0
print("cpu (oracle port, 1 thread): %.0f ms for %d cells = %.4f Gcells/s; GPU result identical: %s" % (dt * 1e3, len(c), len(c) / dt / 1e9, cells.cpu().numpy().tobytes() == c.tobytes()))
