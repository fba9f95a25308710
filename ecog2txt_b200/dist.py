"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the single
gradient all-reduce per step; the gradient bucket is the library's own flat fp32 buffer, aliased
zero-copy as a torch tensor.  (The reference hard-codes training_GPUs=[0],
/root/reference/ecog2txt/trainers.py:131.)"""
from __future__ import annotations

from . import _lib as L


class _DevBuf:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def flat_tensor(engine, which: int = L.GRAD):
    """torch view (no copy) of one of the engine's flat parameter-sized buffers (device memory; host memory
    only for the test-only kernel-emulation build, whose "device" buffers are malloc'ed)."""
    import torch
    ptr, n = engine.flat_buffer(which)
    if getattr(engine, "emulated", False):
        import ctypes
        import numpy as np
        return torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(n,)))
    return torch.as_tensor(_DevBuf(ptr, n), device=torch.device("cuda", engine.cfg.device))


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_grads(engine, grads_tensor, ntok_local: float):
    """One flat all-reduce (sum) of the gradient bucket + the token count; returns the global count."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ntok_local
    t = torch.tensor([ntok_local], dtype=torch.float64, device=grads_tensor.device)
    dist.all_reduce(grads_tensor)
    dist.all_reduce(t)
    return float(t.item())
