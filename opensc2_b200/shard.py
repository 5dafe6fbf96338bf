"""Batch sharding over the GPUs of one box: conductors of a sweep never
exchange data (the reference raises on thermally coupled conductors, simulation.py:281-303),
so rank g of G owns a contiguous slice of the batch, steps it with its own handle, and the only
collective is the final gather of results (`torch.distributed`, NCCL on GPUs / gloo in the
CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of `total` conductors for `rank` (sizes differ by <= 1)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} of {world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(total: int, world: int) -> List[int]:
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def gather_to_all(local: np.ndarray, total: int, device=None) -> np.ndarray:
    """Final result gather: every rank contributes its [n_local, ...] slice (shard_range order) and
    receives the full [total, ...] array.  Uneven shards are padded to the largest one."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank}: {local.shape[0]} rows, expected {sizes[rank]}")
    n_max = max(sizes)
    pad = np.zeros((n_max,) + local.shape[1:], dtype=local.dtype)
    pad[:local.shape[0]] = local
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous())          # concatenation along the leading axis
    out = out.cpu().numpy().reshape((world, n_max) + tuple(t.shape[1:]))
    return np.concatenate([out[r, :sizes[r]] for r in range(world)], axis=0)


def any_rank_failed(status: np.ndarray, device=None) -> bool:
    """all_reduce(MAX) of 'some conductor of my shard is singular / out of range'."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([int(np.any(status != 0))], dtype=torch.int32)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(int(t.item()))
