"""Source sharding over ranks and the phi_ion reduction (reference: pyc2ray/evolve.py:360-373,433-437).

One process per GPU.  The ray trace is embarrassingly parallel over sources; the only exchange on the
path is the sum of the per-rank rate grids, done as one in-place all-reduce (NCCL over NVLink on
GPUs; gloo on CPU tensors in the host-logic tests).  Every rank then runs the (cheap, deterministic)
chemistry pass on the full grid, which replaces the reference's rank-0 chemistry + four N^3 broadcasts
(evolve.py:439-497) with zero further traffic.
"""
import numpy as np

__all__ = ["shard_bounds", "allreduce_sum_", "DeviceView", "device_tensor"]


def shard_bounds(NumSrc, rank, nprocs):
    """Contiguous block split, remainder to the last rank (evolve.py:362-367)."""
    perrank = NumSrc // nprocs
    i_start = int(rank * perrank)
    i_end = int((rank + 1) * perrank) if rank != nprocs - 1 else NumSrc
    return i_start, i_end


def allreduce_sum_(tensor, group=None):
    """In-place sum over ranks of a torch tensor (CUDA -> NCCL, CPU -> gloo)."""
    import torch.distributed as dist
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


class DeviceView:
    """Zero-copy ``__cuda_array_interface__`` view of a context-owned device buffer."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def device_tensor(ptr, n):
    """torch.float64 CUDA tensor aliasing ``n`` doubles at device address ``ptr``."""
    import torch
    return torch.as_tensor(DeviceView(ptr, n), device="cuda")
