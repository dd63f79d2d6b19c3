"""Data-parallel plumbing: the path shards over the batch.  Rank r owns a contiguous block
of sample rows; parameters and R-vectors are replicated; after the local step the parameter gradients
and R-gradients are summed across ranks -- the multi-device form of Gradient.Add (gradient.go:28-30,
108-120; the reference's gradient is a SUM over samples, never a mean).

On GPUs the collective lives BEHIND the C ABI (af_dp_init / af_allreduce_sum / af_mlp_step_dp, csrc/dp.cu: NCCL over
NVLink, resolved inside the library); the host side only has to ship the 128-byte communicator id from rank 0 to the
other ranks -- dp_unique_id() + any broadcast (torch.distributed in bench.py, a file or socket elsewhere).  The torch
helpers below (allreduce_sum_, plan_grad_slab) are the gloo path of the CPU tests and the round-1 A/B transport."""
from __future__ import annotations

import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def dp_unique_id() -> bytes:
    """Rank 0's half of the rendezvous: af_dp_unique_id (ncclGetUniqueId behind the ABI)."""
    import ctypes as C

    from . import _lib
    buf = C.create_string_buffer(_lib.AF_DP_ID_BYTES)
    _lib.check(_lib.load().af_dp_unique_id(buf))
    return buf.raw


def dp_init_torch(ctx, device_index: int):
    """Create ctx's communicator with torch.distributed (already initialised, any backend) as the id courier."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = f"cuda:{device_index}" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(dp_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    ctx.dp_init(world, rank, bytes(t.cpu().numpy().tobytes()))


def dp_init_local(contexts) -> None:
    """One process, several devices: af_dp_init_local (ncclCommInitAll) over one Context per device."""
    import ctypes as C

    from . import _lib
    arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
    _lib.check(_lib.load().af_dp_init_local(arr, len(contexts)))


def shard_rows(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """(first row, row count) of rank's block; blocks differ by at most one row and cover [0, n)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_total, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


class CudaArrayView:
    """Zero-copy torch view of a device buffer owned by libautofunc_b200 (__cuda_array_interface__)."""

    def __init__(self, ptr: int, n_doubles: int):
        self.__cuda_array_interface__ = {"shape": (n_doubles,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def allreduce_sum_(tensor, group=None):
    """In-place sum all-reduce of a gradient slab (any torch tensor: CUDA -> NCCL, CPU -> gloo)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def plan_grad_slab(plan, device_index: int):
    """The plan's contiguous {grad | rgrad} block as one torch CUDA tensor (one collective per step)."""
    import torch
    g0 = plan.grad(0)
    n = (plan.input().ptr - g0.ptr) // 8
    return torch.as_tensor(CudaArrayView(g0.ptr, n), device=f"cuda:{device_index}")
