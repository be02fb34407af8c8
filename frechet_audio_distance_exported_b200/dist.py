"""Multi-GPU plumbing of the FAD path: clips shard across ranks with NO data-path
collective; the only exchange is one all-reduce (sum, fp64) of the packed sufficient statistics
{n, sum x, sum x x^T} of both sets.  Works with the NCCL backend on GPUs and with gloo on CPU
tensors (used by the world_size-2 CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `n_items` owned by `rank`; blocks differ by at most one item and
    tile the range exactly (so results do not depend on `world`)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(int(n_items), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def world_info(group=None) -> Tuple[int, int]:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allreduce_acc(acc: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks of the packed fp64 statistics buffer; no-op for a single process."""
    import torch.distributed as dist
    if acc.dtype != torch.float64:
        raise TypeError("statistics are exchanged in fp64")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc
