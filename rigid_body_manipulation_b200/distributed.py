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
    """In-place sum of the [Y^T Y | Y^T f | f^T f | n] pack over all ranks (no-op without an initialised group).
    A (k, 112) tensor holds one pack per object and is reduced in ONE collective (grouped all-reduce)."""
    if pack.dim() not in (1, 2) or pack.shape[-1] != 112 or pack.dtype != torch.float64 or not pack.is_contiguous():
        raise ValueError("pack must be a contiguous float64 tensor of shape (112,) or (k, 112)")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=group)
    return pack


class NcclGramReducer:
    """The Gram all-reduce through the library's own C entry point (rbm_allreduce_gram) on a communicator created from a
    unique id that is broadcast with torch.distributed.  Equivalent to `allreduce_gram`; exists so that applications
    binding the C ABI directly (no torch collectives) have the complete path.  Must be constructed collectively."""

    def __init__(self, device: int, group=None):
        import ctypes as C

        from . import _lib

        self._lib = _lib.load()
        self._check = _lib.check
        if not self._lib.rbm_nccl_available():
            raise _lib.RbmNcclError(_lib.last_error())
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            self._check(self._lib.rbm_nccl_unique_id(buf), "rbm_nccl_unique_id")
        box = [bytes(buf)]
        dist.broadcast_object_list(box, src=0, group=group)
        idbuf = (C.c_ubyte * 128).from_buffer_copy(box[0])
        comm = C.c_void_p()
        self._check(self._lib.rbm_nccl_comm_create(idbuf, world, rank, int(device), C.byref(comm)), "rbm_nccl_comm_create")
        self._comm = comm
        self.device = int(device)

    def __call__(self, pack: torch.Tensor) -> torch.Tensor:
        import ctypes as C

        if pack.dim() not in (1, 2) or pack.shape[-1] != 112 or pack.dtype != torch.float64 or not pack.is_cuda or not pack.is_contiguous():
            raise ValueError("pack must be a contiguous CUDA float64 tensor of shape (112,) or (k, 112)")
        stream = C.c_void_p(torch.cuda.current_stream(pack.device).cuda_stream)
        count = pack.numel() // 112
        self._check(self._lib.rbm_allreduce_gram_n(self._comm, C.c_void_p(pack.data_ptr()), count, stream), "rbm_allreduce_gram_n")
        return pack

    def close(self):
        if getattr(self, "_comm", None):
            self._lib.rbm_nccl_comm_destroy(self._comm)
            self._comm = None


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Max of a host scalar over ranks (timing convention: device time, max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
