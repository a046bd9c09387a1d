"""Host-side sharding of the red-noise path across ranks (one process per GPU).

The path has no exchange step: every sample is a pure function of (seed, params, lattice
index), and a window's dependency cone is recomputed, never communicated (rn_tile.cuh).  So
  * one large field shards by contiguous row slabs (rn_window per rank, seam-exact), and
  * a batch of independent fields (streaming terrain chunks, distinct seeds) shards by field.
`all_gather_slabs` is the optional final assembly of a slab-sharded field on every rank
(NCCL over NVLink on GPUs; gloo in the CPU tests) -- the only collective of the whole path.
"""
from __future__ import annotations

from typing import List, Tuple


def slab_rows(side: int, world: int) -> int:
    """Rows per rank: equal slabs (the all-gather needs equal shapes); the last rank may own fewer."""
    return (side + world - 1) // world


def slab_window(side: int, rank: int, world: int) -> Tuple[int, int, int, int]:
    """(i0, j0, ni, nj) of rank's row slab of a side x side field; ni == 0 for a rank past the end."""
    rows = slab_rows(side, world)
    i0 = min(rank * rows, side)
    ni = max(0, min(rows, side - i0))
    return i0, 0, ni, side


def batch_shard(ntiles: int, rank: int, world: int) -> List[int]:
    """Indices of the fields of a batch this rank generates (round-robin)."""
    return list(range(rank, ntiles, world))


def all_gather_slabs(local, side: int, group=None):
    """Assembles the full (side, side, 4) field on every rank from equal-height row slabs.

    `local` is this rank's (slab_rows, side, 4) float32 tensor (rows past the end of the field are
    padding).  One all_gather_into_tensor of contiguous blocks -- NCCL over NVLink/NVSwitch when
    the tensors live on GPUs.  Returns the (side, side, 4) tensor (a view of the gathered buffer).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rows = slab_rows(side, world)
    assert local.shape == (rows, side, 4) and local.dtype == torch.float32 and local.is_contiguous()
    out = torch.empty((world * rows, side, 4), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out[:side]
