"""Host ceiling of the end-to-end path: bare device-to-host copies of one depth-12 field (268.6 MB) into PINNED host memory on
1 .. N GPUs AT THE SAME TIME (one process per GPU), no synthesis at all.  What bench.py's `e2e` can at best reach at N GPUs is
N fields per (this time); if the aggregate stops growing with N the limiter is the host (PCIe switches / root complex / memory),
not the library.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/micro/d2h_concurrent.py

Also times a pageable destination (what `new TerrainElement[]`, Terrain.cpp:45-49, gives an unmodified caller) on rank 0 alone."""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S = 4097
dev = torch.rand((S, S, 4), device="cuda", dtype=torch.float32)
host = torch.empty((S, S, 4), dtype=torch.float32, pin_memory=True)
nbytes = dev.numel() * 4


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(dst, reps):
    for _ in range(2):
        dst.copy_(dev, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


dt = timed(host, 10)
out = {"n_gpus": world, "bytes_per_gpu": nbytes, "pinned_ms": dt * 1e3, "pinned_gbs_per_gpu": nbytes / dt / 1e9,
       "pinned_gbs_aggregate": world * nbytes / dt / 1e9, "fields_per_s_ceiling": world / dt,
       "gsamples_per_s_ceiling": world * S * S / dt / 1e9}
if world == 1:
    pageable = torch.empty((S, S, 4), dtype=torch.float32)
    dtp = timed(pageable, 5)
    out.update(pageable_ms=dtp * 1e3, pageable_gbs=nbytes / dtp / 1e9)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
