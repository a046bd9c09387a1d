#!/usr/bin/env python
"""bench.py -- red-noise field synthesis throughput (Gsamples/s) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one unit of work per GPU: by default one
Terrain(depth=12, size=500, amplitude=0.2) field = 4097x4097 samples of 16 B (BASELINE.json
configs[1]).  With N > 1 GPUs every rank synthesises its own field from its own seed
(independent units, no data-path collective): weak scaling.  `--workload slab` instead shards
ONE depth-14 field by row slabs across the ranks (strong scaling).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ELEM_BYTES = 16  # sizeof(Terrain::TerrainElement), Terrain.h:15-19: ALGORITHMIC bytes written per sample
BASE_SEED = 12345


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)", d
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", {}


NCU_METRICS = os.path.join("profiles", "r02_march_kernel_ncu_metrics.json")  # written by profiles/micro/ncu_export.py from the
# `ncu --set full` capture of THIS command's output kernel (k_march, 2049^2 -> 4097^2); refreshed whenever the kernel changes


NCU_METRICS_VOLUME = os.path.join("profiles", "r02_vol_rows_kernel_ncu_metrics.json")  # the same for `--workload volume` (k_vol_rows, 257^3 -> 513^3)


def _ncu_file(args):
    if args.workload == "field" and args.depth == 12 and args.algo in (0, 2):
        return NCU_METRICS
    if args.workload == "volume" and args.depth == 9:
        return NCU_METRICS_VOLUME
    return None


def _ncu_metrics(args):
    try:
        return json.load(open(os.path.join(ROOT, _ncu_file(args))))
    except Exception:
        return None


def _mbytes(entry):
    v, unit = float(entry[0]), entry[1].lower()
    return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]


def alu_roofline(args, value_gs, kernel_ms, my_samples, sm_mhz):
    """The instruction-issue ("ALU") bound of the output kernel: thread instructions per sample counted by ncu on the
    committed capture of this command (NCU_METRICS) against 148 SMs x 4 schedulers x 32 lanes x the SM clock seen during this
    run.  The bit-exact contract forbids FMA contraction, so this bound sits near the HBM-write bound."""
    d = _ncu_metrics(args)
    if not d:
        return None
    try:
        warp_inst = float(d["smsp__inst_executed.sum"][0])
        lanes = float(d["smsp__thread_inst_executed_per_inst_executed.ratio"][0])
        side = (1 << args.depth) + 1
        inst_per_sample = warp_inst * lanes / (side ** 3 if args.workload == "volume" else side * side)
        mhz = float(sm_mhz or 1965.0)
        peak = 148 * 4 * 32 * mhz * 1e6           # thread instructions / s
        bound = peak / inst_per_sample / 1e9      # Gsamples/s of the output kernel at 100 % issue
        achieved = my_samples / (kernel_ms * 1e-3) / 1e9
        return {"bound": "issue", "thread_inst_per_sample": inst_per_sample, "peak_thread_inst_per_s": peak,
                "gsamples_bound": bound, "achieved_gsamples_kernel": achieved, "frac": achieved / bound,
                "ncu": {"issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
                        "l1tex_lsu_wavefronts_pct": float(d["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"][0]),
                        "pipe_fma_pct": float(d["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"][0]),
                        "pipe_alu_pct": float(d["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"][0])},
                "source": _ncu_file(args) + " (ncu --set full, " + d["kernel"][0] + ")"}
    except Exception:
        return None


def ncu_traffic(args):
    """dram__bytes_read.sum + dram__bytes_write.sum of the output-stage kernel per launch, from the committed
    `ncu --set full` capture of this command (NCU_METRICS); None for other workloads."""
    d = _ncu_metrics(args)
    try:
        return int(_mbytes(d["dram__bytes_read.sum"]) + _mbytes(d["dram__bytes_write.sum"])) if d else None
    except Exception:
        return None


# --------------------------------------------------------------------------- clocks ---
class ClockSampler:
    """SM clock + throttle reasons sampled through NVML every ~2 ms by a background thread while the GPU works
    (nvidia-smi -lms cannot resolve a timed region of a few milliseconds); falls back to nvidia-smi."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4),
               ("hw_power_brake_slowdown", 0x80))

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = False
        self._thr = None
        self._nvml = None

    def start(self):
        import threading
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES if it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self._nvml = (pynvml, h)
        except Exception:
            self._nvml = None
            return

        def loop():
            pynvml, h = self._nvml
            while not self._stop:
                try:
                    self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    for name, bit in self.REASONS:
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                time.sleep(0.0002)

        self._thr = threading.Thread(target=loop, daemon=True)
        self._thr.start()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": "nvml"}
        if self._thr is not None:
            self._stop = True
            self._thr.join(timeout=2)
        if self.samples:
            out.update(sm_mhz=statistics.median(self.samples), reasons=sorted(self.reasons), samples=len(self.samples))
            return out
        try:  # fallback: one nvidia-smi query
            q = "clocks.sm,clocks.max.sm"
            r = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            out.update(sm_mhz=float(r[0]), sm_max_mhz=float(r[1]), samples=1, source="nvidia-smi")
        except Exception:
            pass
        return out


# ------------------------------------------------------------------- reference arm ---
def _ref_worker(args):
    depth, seed = args
    from tests import oracle_lib
    ref = oracle_lib.load_ref()
    t0 = time.perf_counter()
    _, calls, times = ref.terrain2d(depth, seed, volcano=1, want_times=True)
    return time.perf_counter() - t0, times


def _port_worker(args):
    depth, seed = args
    from tests import oracle_lib
    o = oracle_lib.load_oracle()
    t0 = time.perf_counter()
    o.terrain2d(oracle_lib.params(seed=seed, depth=depth, apply_volcano=1))
    dt = time.perf_counter() - t0
    return dt, (0.0, dt, 0.0)


def cpu_fields_per_second(depth: int, procs: int, rounds: int, budget_s: float):
    """Times the reference's own CPU generator (oracle/_ref, else the oracle port): `procs`
    independent fields (distinct seeds) generated concurrently, one process each (the reference
    is single-threaded with global rand() state), `rounds` times or until the budget is spent.
    Returns (fields/s aggregate, kind, per-round seconds, pure-pass seconds of one field)."""
    from concurrent.futures import ProcessPoolExecutor
    from tests import oracle_lib
    kind = "reference" if oracle_lib.load_ref() is not None else "port"
    worker = _ref_worker if kind == "reference" else _port_worker
    if kind == "port":
        oracle_lib.load_oracle()
    round_s, pass_s = [], []
    t_start = time.perf_counter()
    with ProcessPoolExecutor(max_workers=procs) as ex:
        list(ex.map(worker, [(min(depth, 8), 1)] * procs))  # warm the pool / page in the library
        for r in range(rounds):
            t0 = time.perf_counter()
            res = list(ex.map(worker, [(depth, BASE_SEED + r * procs + q) for q in range(procs)]))
            round_s.append(time.perf_counter() - t0)
            pass_s.extend(t[1] for _, t in res)
            if time.perf_counter() - t_start > budget_s:
                break
    fps = procs / statistics.median(round_s)
    return fps, kind, round_s, statistics.median(pass_s)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    depth = args.depth
    side = (1 << depth) + 1
    if args.workload == "volume":
        # the 3D volume is an extension without a reference implementation: the arm times its CPU specification
        from tests import oracle_lib
        o = oracle_lib.load_oracle()
        d_cpu = min(depth, 7)
        n_cpu = ((1 << d_cpu) + 1) ** 3
        times = []
        for _ in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            o.volume3d(BASE_SEED, d_cpu, 0.3)
            times.append(time.perf_counter() - t0)
        timed = times[args.warmup:] or times
        ms = 1e3 * statistics.mean(timed)
        value = n_cpu / (ms * 1e-3) / 1e9
        print(json.dumps({
            "impl": "reference", "metric": "red-noise field synthesis throughput", "value": value, "unit": "Gsamples/s", "n_gpus": args.gpus,
            "steps": len(timed), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": 1, "kind": "port",
                             "sample": f"one {(1 << d_cpu) + 1}^3 volume per step by the CPU specification oracle/rn_oracle3d.c (no reference "
                                       "implementation of a volume generator exists)", "host_cpus": os.cpu_count()},
            "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return 0
    samples = side * side
    procs = max(1, min(os.cpu_count() or 1, args.cpu_procs or (os.cpu_count() or 1)))
    fps, kind, round_s, pass_s = cpu_fields_per_second(depth, procs, args.warmup + args.steps, budget_s=240.0)
    timed = round_s[args.warmup:] if len(round_s) > args.warmup else round_s
    ms = 1e3 * statistics.mean(timed)
    value = procs * samples / (ms * 1e-3) / 1e9
    line = {
        "impl": "reference", "metric": "red-noise field synthesis throughput", "value": value, "unit": "Gsamples/s",
        "n_gpus": args.gpus, "steps": len(timed), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": procs, "kind": kind,
                         "sample": f"{procs} independent Terrain(depth={depth}) fields per step, one single-threaded process each "
                                   f"(grid init + diamond-square passes + addVolcano); pure passes of one field: {pass_s * 1e3:.0f} ms",
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args):
    side = (1 << args.depth) + 1
    what = {
        "field": f"{side}x{side} 2D red-noise heightfield per rank (Terrain depth={args.depth}: {args.depth - 2} displacement levels/'octaves'), "
                 f"amplitude 0.2, size 500, seed {BASE_SEED}+rank, addVolcano on",
        "slab": f"row slabs of ONE {side}x{side} field (depth={args.depth}, seed {BASE_SEED}) sharded over the ranks"
                + (", assembled on every rank by one NCCL all-gather" if getattr(args, "gather", False) else ""),
        "batch": f"batch of {getattr(args, 'tiles', 0)} independent {side}x{side} fields (depth={args.depth}, seeds {BASE_SEED}+k), round-robin over the ranks",
        "volume": f"{side}^3 3D density volume ({args.depth - 1} displacement levels), seed {BASE_SEED}, amplitude 0.3, i-slabs over the ranks -- "
                  "EXTENSION defined by oracle/rn_oracle3d.c, the reference has no volume generator",
    }[args.workload]
    par = {"field": f"dp{args.gpus}: one process per GPU, one independent field each, no collective",
           "slab": f"slab{args.gpus}: equal row slabs, dependency cone recomputed per rank, no data-path collective",
           "batch": f"dp{args.gpus}: fields round-robin over the ranks, no collective",
           "volume": f"slab{args.gpus}: i-slabs of one volume, dependency cone recomputed per rank, no collective"}[args.workload]
    return {
        "workload": what,
        "baseline_config": {"field": "configs[1]: 4096x4096 2D heightfield, 12 octaves, single B200",
                            "slab": "2D analogue of configs[3] (slab-sharded large field; the reference has no 3D generator)",
                            "batch": "configs[4]: batch of independent 1024x1024 tiles with distinct seeds",
                            "volume": {9: "configs[2]: 512^3 3D density volume, 8 octaves, single B200",
                                       11: "configs[3]: 2048^3 3D volume, 10 octaves, slab-sharded"}.get(args.depth, f"3D depth {args.depth}")}[args.workload]
                           if args.depth in (9, 10, 11, 12, 14) else f"depth {args.depth}",
        "fields_per_step_per_rank": 1 if args.workload != "batch" else None,
        "samples_per_field": side * side if args.workload != "volume" else side ** 3,
        "bytes_per_sample": ELEM_BYTES if args.workload != "volume" else 4,
        "cache_policy": "the output of a step (268 MB at depth 12) is larger than the 126 MB L2 and there is no input; no explicit L2 flush",
        "parallelism": par,
    }


def final_kernel_name(args):
    """The kernel that writes the output of this workload (rn_plan.h: plan_stages)."""
    d = args.depth
    if args.workload == "volume":
        half, full = (1 << (d - 1)) + 1, (1 << d) + 1
        if d >= 7:
            return (f"k_vol_rows<4,4,8> (3D extension, warp-row tile kernel, last level {half}^3 -> {full}^3: cube centres, face centres, edge "
                    "midpoints of a 4 x 4 x 64-cell tile in shared memory, 32-bit lattice hash, packed FP32 pairs, full-row stores)")
        return f"k_vol_level (3D extension, thread-per-cell tile kernel, last level {half}^3 -> {full}^3)"
    half, full = (1 << (d - 1)) + 1, (1 << d) + 1
    if d >= 10:
        if args.algo == 3:
            return f"k_stream<128,8> (round-1 block streaming kernel, last level {half}^2 -> {full}^2)"
        return (f"k_march (warp-synchronous march, last level {half}^2 -> {full}^2: averageSquare/generateSquare, both generateDiamond loops, "
                "ordered shade gather, addVolcano, 256-bit sector stores)")
    return f"k_stage<2,64,512> (fused tile kernel, last two levels -> {full}^2)"


def strong_scaling_records(args, ctx, capi, sharding, torch, dist, rank, world, stream, barrier):
    """north_star's scaling configs in front of the driver: ONE depth-14 field (16385^2) in row slabs and ONE 2049^3 density
    volume in i-slabs over the ranks of this run -- fixed total work, no data-path collective (each rank recomputes its slab's
    dependency cone) -- each beside the same job on rank 0 alone, measured in the same run, so that efficiency =
    value / (n_gpus * one_gpu.value) can be computed from this line alone.  Timed like the main figure: barrier +
    synchronize on both sides, CUDA events, max over ranks."""
    out = {}

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0])

    def alone(fn, steps, warm):  # rank 0 only; the other ranks wait at the barriers inside timed()
        return timed(fn if rank == 0 else (lambda: None), steps, warm)

    # (i) one depth-14 field in row slabs
    depth = 14
    side = (1 << depth) + 1
    p = capi.default_params(seed=BASE_SEED, depth=depth, apply_volcano=1)
    i0, j0, ni, nj = sharding.slab_window(side, rank, world)
    dst = torch.empty((max(ni, 1), nj, 4), device="cuda", dtype=torch.float32)
    win = capi.Window(i0, j0, max(ni, 1), nj)
    ms = timed(lambda: ctx.terrain2d_ptr(p, dst.data_ptr(), window=win, dst_is_device=True, stream=stream) if ni else None, 5, 2)
    rec = {"workload": f"ONE {side}x{side} field (depth 14) in {world} row slab(s), dependency cone recomputed per rank, no collective",
           "samples": side * side, "n_gpus": world, "ms_per_step": ms, "value": side * side / (ms * 1e-3) / 1e9, "unit": "Gsamples/s"}
    del dst
    if world > 1:
        full = torch.empty((side, side, 4), device="cuda", dtype=torch.float32) if rank == 0 else None
        ms1 = alone(lambda: ctx.terrain2d_ptr(p, full.data_ptr(), dst_is_device=True, stream=stream), 5, 2)
        rec["one_gpu"] = {"ms_per_step": ms1, "value": side * side / (ms1 * 1e-3) / 1e9}
        del full
    else:
        rec["one_gpu"] = {"ms_per_step": ms, "value": rec["value"]}
    rec["efficiency_inputs"] = "value / (n_gpus * one_gpu.value)"
    out["field_depth14_row_slabs"] = rec
    torch.cuda.empty_cache()

    # (ii) one 2049^3 volume in i-slabs (3D extension, defined by oracle/rn_oracle3d.c; the reference has no volume generator)
    vdepth = 11
    S3 = (1 << vdepth) + 1
    i0, _, ni, _ = sharding.slab_window(S3, rank, world)
    try:
        vol = torch.empty((max(ni, 1), S3, S3), device="cuda", dtype=torch.float32)
        ms = timed(lambda: ctx.volume3d_ptr(BASE_SEED, vdepth, 0.3, vol.data_ptr(), i0=i0, ni=ni, stream=stream) if ni else None, 3, 1)
        rec = {"workload": f"ONE {S3}^3 density volume (3D extension, 10 levels) in {world} i-slab(s) of {sharding.slab_rows(S3, world)} planes, "
                           "slab-sized destinations, dependency cone recomputed per rank, no collective",
               "samples": S3 ** 3, "n_gpus": world, "ms_per_step": ms, "value": S3 ** 3 / (ms * 1e-3) / 1e9, "unit": "Gsamples/s"}
        del vol
        torch.cuda.empty_cache()
        if world > 1:
            full = torch.empty((S3, S3, S3), device="cuda", dtype=torch.float32) if rank == 0 else None
            ms1 = alone(lambda: ctx.volume3d_ptr(BASE_SEED, vdepth, 0.3, full.data_ptr(), stream=stream), 3, 1)
            rec["one_gpu"] = {"ms_per_step": ms1, "value": S3 ** 3 / (ms1 * 1e-3) / 1e9}
            del full
        else:
            rec["one_gpu"] = {"ms_per_step": ms, "value": rec["value"]}
        rec["efficiency_inputs"] = "value / (n_gpus * one_gpu.value)"
        out["volume_2049_i_slabs"] = rec
    except Exception as e:  # e.g. not enough free HBM for the 34.4 GB single-GPU volume on a shared box
        out["volume_2049_i_slabs"] = {"error": repr(e)}
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------ our arm ---
def run_ours(args):
    import torch
    import torch.distributed as dist
    from rednoise_b200 import build, capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N > 1 launch with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # one disjoint set of host cores per rank: a timed region is a few milliseconds of asynchronous enqueues (40 us of host time
        # per 120 us step), and a rank whose enqueue thread is descheduled for a millisecond starves its GPU (max over ranks)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = len(cores) // world
            if per >= 2:
                os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]))
        except (AttributeError, OSError):
            pass
    build.build()  # no-op when the in-tree .so is fresh
    ctx = capi.Context(local_rank)
    ctx.set_algorithm(args.algo)
    if args.no_pipelining:
        ctx.set_pipelining(False)

    depth = args.depth
    side = (1 << depth) + 1
    # a real (non-default) stream: the C ABI treats stream == NULL as "blocking call", and the timed
    # loop must enqueue asynchronously like a production caller would
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    from rednoise_b200 import sharding
    gathered = None
    if args.workload == "field":
        p = capi.default_params(seed=BASE_SEED + rank, depth=depth, apply_volcano=1)
        win = None
        ni, nj = side, side
        my_samples = ni * nj
        dst = torch.empty((ni, nj, 4), device="cuda", dtype=torch.float32)

        def step():
            ctx.terrain2d_ptr(p, dst.data_ptr(), window=win, dst_is_device=True, stream=stream)
    elif args.workload == "slab":  # one field, equal row slabs, halo recomputed per rank, optional NCCL assembly
        p = capi.default_params(seed=BASE_SEED, depth=depth, apply_volcano=1)
        rows = sharding.slab_rows(side, world)
        i0, j0, ni, nj = sharding.slab_window(side, rank, world)
        win = capi.Window(i0, j0, max(ni, 1), nj)
        my_samples = ni * nj
        dst = torch.zeros((rows, nj, 4), device="cuda", dtype=torch.float32)

        def step():
            if ni:
                ctx.terrain2d_ptr(p, dst.data_ptr(), window=win, dst_is_device=True, stream=stream)
            if args.gather and world > 1:
                sharding.all_gather_slabs(dst, side)
    elif args.workload == "volume":  # 3D extension: i-slabs of ONE (W+1)^3 density volume, halo recomputed per rank
        S3 = side
        i0, _, ni, _ = sharding.slab_window(S3, rank, world)
        my_samples = ni * S3 * S3
        nj = S3
        vol = torch.empty((max(ni, 1), S3, S3), device="cuda", dtype=torch.float32)  # slab-sized: only this rank's planes
        p = None
        win = None
        seeds = []

        def step():
            if ni:
                ctx.volume3d_ptr(BASE_SEED, depth, 0.3, vol.data_ptr(), i0=i0, ni=ni, stream=stream)
    else:  # batch: args.tiles independent fields with distinct seeds, round-robin over the ranks
        p = capi.default_params(seed=BASE_SEED, depth=depth, apply_volcano=1)
        win = None
        mine = sharding.batch_shard(args.tiles, rank, world)
        seeds = [BASE_SEED + k for k in mine]
        ni, nj = side, side
        my_samples = len(mine) * side * side
        dst = torch.empty((max(len(mine), 1), side, side, 4), device="cuda", dtype=torch.float32)

        def step():
            if seeds:
                ctx.terrain2d_batch_ptr(p, seeds, dst.data_ptr(), dst_is_device=True, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    import gc
    gc.collect()
    gc.disable()  # no collector pause inside the timed region
    barrier()
    ev[0].record()
    t_host0 = time.perf_counter()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    host_enqueue_us = (time.perf_counter() - t_host0) / args.steps * 1e6  # host time per step of the asynchronous calls
    barrier()
    gc.enable()
    clocks = sampler.stop() if rank == 0 else None  # sampled during the timed region only
    total_ms = ev[0].elapsed_time(ev[-1])
    own_total_ms = total_ms
    per_step = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    launches = ctx.launches - launches0
    # per-launch device time of the stage chain (CUDA events on the launching stream), measured
    # in a separate pass so the event records do not sit inside the timed region above
    stage_ms = None
    prof_samples = my_samples  # samples the profiled launches produce
    if args.algo != 1:
        prof_step = step
        if args.workload == "batch" and seeds:
            # the library cuts a batch into chunks (1 GB of intermediate grids) and reports the launches of the LAST chunk: profile
            # a sub-batch that is one chunk, and take the roofline bytes from it
            prof_tiles = max(1, min(len(seeds), ctx.batch_chunk(p, len(seeds))))
            prof_samples = prof_tiles * side * side

            def prof_step():
                ctx.terrain2d_batch_ptr(p, seeds[:prof_tiles], dst.data_ptr(), dst_is_device=True, stream=stream)
        ctx.set_profiling(True)
        acc = []
        for _ in range(max(5, min(args.steps, 20))):
            prof_step()
            acc.append(ctx.last_stage_times())
        ctx.set_profiling(False)
        if acc and acc[0]:
            ns = len(acc[0])
            mean = [statistics.mean(a[k] for a in acc) for k in range(ns)]
            stage_ms = {"launch_ms": mean, "final_ms": mean[-1], "final_name": final_kernel_name(args),
                        "share_of_chain": mean[-1] / sum(mean)}

    # ---- one call at a time (what Core::init does, Core.cpp:65): no call in flight underneath, so nothing of the stage chain is hidden ----
    single_ms = None
    if args.workload in ("field", "volume"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lat = []
        for _ in range(10):
            torch.cuda.synchronize()
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            lat.append(e0.elapsed_time(e1))
        single_ms = statistics.median(lat)

    # ---- e2e: the same C-ABI call with a HOST destination (pinned), D2H inside the timed region ----
    e2e_steps = max(2, min(args.steps, 5))
    if args.workload == "volume":
        # bounded pinned buffer (<= 1 GiB per rank): the e2e leg generates and ships a sub-slab of ni_e planes
        ni_e = max(0, min(ni, (1 << 30) // (S3 * S3 * 4)))
        host = torch.empty((max(ni_e, 1), S3, S3), dtype=torch.float32, pin_memory=True)
        e2e_samples = ni_e * S3 * S3

        def e2e_step():
            if ni_e:
                ctx.volume3d_to_host_ptr(BASE_SEED, depth, 0.3, host.data_ptr(), i0=i0, ni=ni_e)  # rn_volume3d_host: blocking, chunked synthesis + D2H
                return float(host[ni_e // 2, S3 // 2, S3 // 2])
            return 0.0
    elif args.workload == "batch":
        e2e_tiles = min(len(seeds), 32)  # bounded pinned buffer; the metric is a rate
        host = torch.empty((max(e2e_tiles, 1), side, side, 4), dtype=torch.float32, pin_memory=True)
        e2e_samples = e2e_tiles * side * side

        def e2e_step():
            if e2e_tiles:
                ctx.terrain2d_batch_ptr(p, seeds[:e2e_tiles], host.data_ptr(), dst_is_device=False)
                return float(host[e2e_tiles - 1, side // 2, side // 2, 1])
            return 0.0
    else:
        host = torch.empty((max(ni, 1), nj, 4), dtype=torch.float32, pin_memory=True)
        e2e_samples = my_samples

        def e2e_step():
            if ni:
                ctx.terrain2d_ptr(p, host.data_ptr(), window=win, dst_is_device=False)  # blocking: synth + D2H + sync
                return float(host[ni // 2, nj // 2, 1])  # the host reads the step's result
            return 0.0
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    # the same call with a PAGEABLE destination (what `new TerrainElement[]`, Terrain.cpp:45-49, gives an unmodified caller)
    e2e_pageable_s = None
    if args.workload == "field" and ni:
        pageable = torch.empty((ni, nj, 4), dtype=torch.float32)
        ctx.terrain2d_ptr(p, pageable.data_ptr(), window=win, dst_is_device=False)
        t0 = time.perf_counter()
        for _ in range(2):
            ctx.terrain2d_ptr(p, pageable.data_ptr(), window=win, dst_is_device=False)
        e2e_pageable_s = (time.perf_counter() - t0) / 2
        del pageable
    strong = None
    if args.workload == "field" and not args.no_strong:
        strong = strong_scaling_records(args, ctx, capi, sharding, torch, dist, rank, world, stream, barrier)
    t = torch.tensor([total_ms, e2e_s, float(my_samples), float(e2e_samples)], device="cuda", dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_s, all_samples, all_e2e_samples = float(tmax[0]), float(tmax[1]), float(tsum[2]), float(tsum[3])
        mine = torch.tensor([own_total_ms / args.steps, host_enqueue_us], device="cuda", dtype=torch.float64)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        per_rank = {"ms_per_step": [round(float(x[0]), 5) for x in every], "host_enqueue_us_per_step": [round(float(x[1]), 1) for x in every]}
    else:
        per_rank = {"ms_per_step": [round(own_total_ms / args.steps, 5)], "host_enqueue_us_per_step": [round(host_enqueue_us, 1)]}
        all_samples, all_e2e_samples = float(my_samples), float(e2e_samples)

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = all_samples / (ms_per_step * 1e-3) / 1e9
        peak, peak_src, _ = load_peaks()
        # dominant kernel: the final fused stage when the library reports stage times, else the whole step
        if stage_ms and stage_ms.get("final_ms"):
            k_ms, k_name = stage_ms["final_ms"], stage_ms.get("final_name", "final stage")
        else:
            k_ms, k_name = statistics.mean(per_step), "whole step (all launches)"
        bytes_per_sample = 4 if args.workload == "volume" else ELEM_BYTES
        k_samples = prof_samples if (stage_ms and stage_ms.get("final_ms")) else my_samples
        achieved = k_samples * bytes_per_sample / (k_ms * 1e-3) / 1e9
        line = {
            "metric": "red-noise field synthesis throughput", "value": value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if args.workload == "field" else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_step": (my_samples * bytes_per_sample / (statistics.mean(per_step) * 1e-3) / 1e9) / peak,
                         "traffic": ncu_traffic(args),
                         "kernel": k_name, "kernel_ms": k_ms, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": k_samples * bytes_per_sample,
                         "samples_per_s_bound_hbm": peak * 1e9 / bytes_per_sample / 1e9},
            "e2e": {"value": all_e2e_samples / e2e_s / 1e9, "unit": "Gsamples/s",
                    "h2d_bytes_per_step": 12 if args.workload == "volume" else 48 + (4 * len(seeds[:32]) if args.workload == "batch" else 0),
                    "d2h_bytes_per_step": e2e_samples * (4 if args.workload == "volume" else ELEM_BYTES), "ms_per_step": e2e_s * 1e3,
                    "note": ("rn_volume3d_host() with a pinned HOST destination: synthesis + D2H of the samples + host read; the only input is the 12-byte rn_volume_params"
                             if args.workload == "volume" else
                             "rn_terrain2d()/rn_terrain2d_batch() with a pinned HOST destination: synthesis + D2H of the samples + host read; "
                             "the only inputs are the 48-byte rn_params (and the seeds of a batch)"),
                    "pageable_destination": None if e2e_pageable_s is None else
                    {"value_rank0": my_samples / e2e_pageable_s / 1e9, "ms_per_step": e2e_pageable_s * 1e3,
                     "note": "same call, destination from plain malloc (rank 0's own figure)"}},
            "gpu_launches": int(launches), "clocks": clocks, "algorithm": args.algo,
            "per_step_ms": {"min": min(per_step), "median": statistics.median(per_step), "max": max(per_step)},
            "per_rank": per_rank,  # each rank's own device time per step (value uses the max) and the host time of one asynchronous call
        }
        if single_ms:
            line["single_call"] = {"ms": single_ms, "value": my_samples / (single_ms * 1e-3) / 1e9, "unit": "Gsamples/s",
                                   "note": "one asynchronous call between two device synchronisations (rank 0): the latency of an isolated call; "
                                           "`value` above is the steady state of back-to-back calls, whose coarse stages overlap the previous call's output kernel"}
        if strong:
            line["strong"] = strong
        if stage_ms:
            line["stage_ms"] = stage_ms
            alu = alu_roofline(args, value, k_ms, k_samples, (clocks or {}).get("sm_mhz"))
            if alu:
                line["roofline_alu"] = alu
        if world == 1 and not args.no_cpu and args.workload == "volume":
            try:
                from tests import oracle_lib
                o = oracle_lib.load_oracle()
                d_cpu = min(depth, 7)
                t0 = time.perf_counter()
                o.volume3d(BASE_SEED, d_cpu, 0.3)
                dt = time.perf_counter() - t0
                n_cpu = ((1 << d_cpu) + 1) ** 3
                line["cpu_baseline"] = {"value": n_cpu / dt / 1e9, "unit": "Gsamples/s", "cores": 1, "kind": "port",
                                        "sample": f"one {(1 << d_cpu) + 1}^3 volume by the CPU specification (oracle/rn_oracle3d.c, the definition of this "
                                                  f"extension; no reference implementation exists) on one host thread: {dt * 1e3:.0f} ms",
                                        "host_cpus": os.cpu_count()}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "Gsamples/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
        elif world == 1 and not args.no_cpu:
            try:
                procs = 1
                fps, kind, round_s, pass_s = cpu_fields_per_second(depth, procs, rounds=8, budget_s=20.0)
                line["cpu_baseline"] = {
                    "value": fps * side * side / 1e9, "unit": "Gsamples/s", "cores": procs, "kind": kind,
                    "sample": f"{len(round_s)} x Terrain(depth={depth}) on one host thread (the reference is single-threaded): "
                              f"grid init + passes + addVolcano; median {statistics.median(round_s) * 1e3:.0f} ms/field, "
                              f"pure passes {pass_s * 1e3:.0f} ms",
                    "host_cpus": os.cpu_count()}
            except Exception as e:  # the bench line must still print
                line["cpu_baseline"] = {"value": None, "unit": "Gsamples/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="field", choices=["field", "slab", "batch", "volume"])
    ap.add_argument("--tiles", type=int, default=1024, help="batch workload: independent fields per step (all ranks together)")
    ap.add_argument("--gather", action="store_true", help="slab workload: assemble the field on every rank with one NCCL all-gather per step")
    ap.add_argument("--depth", type=int, default=None)
    ap.add_argument("--algo", type=int, default=0)
    ap.add_argument("--cpu-procs", type=int, default=0, help="reference arm: concurrent single-threaded generators (0 = all host cores)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling records (one depth-14 field / one 2049^3 volume in slabs)")
    ap.add_argument("--no-pipelining", action="store_true", help="run the coarse stages on the caller's stream (no overlap between calls)")
    args = ap.parse_args()
    if args.depth is None:
        args.depth = {"field": 12, "slab": 14, "batch": 10, "volume": 9}[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
