"""CPU, world_size 2 over gloo: the multi-rank host logic of the path.  Each rank produces its row
slab / its share of a batch independently (no data-path collective) with the host replay of the
tile algorithm (the CUDA kernels' own source, tests/host_replay) and the slabs are assembled with
the one optional collective (all_gather_slabs).  The assembled field must equal the oracle's full
field bit for bit on every rank: slabs are seam-exact."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, depth, seed, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from rednoise_b200 import sharding
    from tests import oracle_lib
    from tests.test_host_replay import Win, SO

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = C.CDLL(SO)
        L.replay_terrain2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        side = (1 << depth) + 1
        p = oracle_lib.params(seed=seed, depth=depth, apply_volcano=1)
        rows = sharding.slab_rows(side, world)
        i0, j0, ni, nj = sharding.slab_window(side, rank, world)
        local = np.zeros((rows, side, 4), np.float32)
        if ni:
            w = Win(i0, j0, ni, nj)
            assert L.replay_terrain2d(C.byref(p), C.byref(w), local.ctypes.data, side, 1) == 0
        full = sharding.all_gather_slabs(torch.from_numpy(local), side).numpy()
        want = oracle_lib.load_oracle().terrain2d(p)
        ok_field = bool(np.array_equal(full.view(np.uint32), want.view(np.uint32)))
        # batch sharding: every field generated exactly once across the ranks
        mine = sharding.batch_shard(7, rank, world)
        allidx = [None] * world
        dist.all_gather_object(allidx, mine)
        ok_batch = sorted(sum(allidx, [])) == list(range(7))
        q.put((rank, ok_field, ok_batch, (i0, ni)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("depth,seed", [(7, 11), (8, 12345)])
def test_two_rank_slabs_assemble_to_the_full_field(depth, seed):
    import torch.multiprocessing as mp
    from tests import test_host_replay  # noqa: F401  (builds the replay library via its fixture path below)
    import subprocess
    if not os.path.exists(test_host_replay.SO):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", test_host_replay.SO,
                        test_host_replay.SRC], check=True)
    from tests import oracle_lib
    oracle_lib.load_oracle()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, depth, seed, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, ok_field, ok_batch, _ in res:
        assert ok_field, f"rank {rank}: assembled field differs from the oracle"
        assert ok_batch


def test_slab_windows_partition_the_field():
    from rednoise_b200 import sharding
    for side in (9, 257, 4097, 16385):
        for world in (1, 2, 3, 4, 8):
            rows = [sharding.slab_window(side, r, world) for r in range(world)]
            assert rows[0][0] == 0 and sum(w[2] for w in rows) == side
            for a, b in zip(rows, rows[1:]):
                assert a[0] + a[2] == b[0] or b[2] == 0
