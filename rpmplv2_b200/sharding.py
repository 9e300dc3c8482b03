"""Sharding of large configuration batches over the GPUs of one box.

Configurations are independent, so the batch is cut into contiguous blocks, one per rank (one process per GPU); the robot and
obstacle tables are replicated. There is no collective on the data path: each rank runs the ordinary engine calls on its block.
The only communication is the (optional) gather of the small per-configuration results to one rank — `gather_rows` below, a
`torch.distributed` gather over NCCL/NVLink on the GPU box (gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import numpy as np


def shard_range(N: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [first, first + count) of rank `rank`: i in [rank*N/world, (rank+1)*N/world)."""
    if not (world >= 1 and 0 <= rank < world and N >= 0):
        raise ValueError("bad shard request")
    first, last = (rank * N) // world, ((rank + 1) * N) // world
    return first, last - first


def gather_rows(local, N: int, rank: int, world: int, dst: int = 0, group=None, device=None):
    """Gathers the per-rank row blocks (shard_range order) of a result array to rank `dst`. `local` is a numpy array or a torch
    tensor whose first dimension is this rank's shard; returns the full [N, ...] array on `dst` (same kind as `local`), None elsewhere.
    world == 1 returns `local` unchanged. Blocks are padded to the largest shard so that one fixed-size gather is enough."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    as_numpy = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if as_numpy else local.contiguous()
    if device is not None:
        t = t.to(device)
    counts = [shard_range(N, r, world)[1] for r in range(world)]
    if t.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {t.shape[0]} rows, expected {counts[rank]}")
    width = max(counts)
    padded = torch.zeros((width,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    padded[: t.shape[0]] = t
    bins = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bins, dst=dst, group=group)
    if rank != dst:
        return None
    full = torch.cat([b[:c] for b, c in zip(bins, counts)], dim=0)
    return full.cpu().numpy() if as_numpy else full


def run_sharded(compute: Callable[[int, int], Sequence], N: int, rank: int, world: int, dst: Optional[int] = 0, group=None, device=None):
    """compute(first, count) -> tuple of arrays with `count` rows each (e.g. d_c, profile of the rank's block, produced with the
    stateless generator or a slice of a host array). Returns the gathered tuple on `dst` (None elsewhere); dst=None skips the gather
    and returns the local block — the mode dataset sweeps use, each rank writing its own shard."""
    first, count = shard_range(N, rank, world)
    local = tuple(compute(first, count))
    if dst is None or world == 1:
        return local
    out = tuple(gather_rows(a, N, rank, world, dst, group, device) for a in local)
    return out if rank == dst else None
