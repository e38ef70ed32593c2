"""Multi-GPU plumbing: one process per GPU, samples sharded contiguously, a collective only where the path has one.

  * inverse dynamics, regressor rows, linearisation: every sample is independent -> NO data-path collective;
    `shard_range` gives each rank its slice.
  * identification: each rank accumulates the Gram pack of its shard (rbm_regressor_gram_*), then ONE sum all-reduce of
    the 112-double pack (`allreduce_gram`, NCCL over NVLink on GPUs, gloo on CPU for tests), after which every rank
    solves the same 10x10 system.  The collective is enqueued on the same stream right behind the finalisation kernel;
    there is no host synchronisation in between.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [start, stop) of rank's samples; the first n_total % world ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    base, rem = divmod(int(n_total), world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_gram(pack: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum of the [Y^T Y | Y^T f | f^T f | n] pack over all ranks (no-op without an initialised group)."""
    if pack.numel() != 112 or pack.dtype != torch.float64:
        raise ValueError("pack must be 112 float64 values")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=group)
    return pack


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Max of a host scalar over ranks (timing convention: device time, max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
