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


def allreduce_step(grads_and_count):
    """THE collective of a data-parallel training step: one flat all-reduce (sum) over `flat_tensor(engine,
    GRAD_AND_COUNT)` -- the 57 MB gradient bucket with the token count riding in its tail.  Enqueue-only: no host
    synchronisation; follow with engine.adam_ema_step_dev(None)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(grads_and_count)


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


class BucketedAllReduce:
    """All-reduce of the gradient buffer bucket by bucket on a side stream, started while the backward pass of the same step
    is still running (the library completes the buffer in flat ranges and records a CUDA event after each:
    e2t_set_grad_buckets / e2t_grad_bucket_*).  The NCCL kernels run on the SMs the persistent recurrent kernels leave idle
    (100 of 148), so only the last, small bucket (layer 0 + the conv) is exposed.

        ar = BucketedAllReduce(engine)            # once; turns bucketing on
        engine.train_step_grads(..., want_loss=False)    # enqueue only, no host sync
        ntok_global = ar.reduce(ntok_device_tensor)       # enqueues the per-bucket all-reduces, joins the streams
        engine.adam_ema_step(1 / ntok_global)
    """

    def __init__(self, engine):
        import torch
        self.engine = engine
        self.grads = flat_tensor(engine, L.GRAD)
        self.emulated = getattr(engine, "emulated", False)
        self.comm = None if self.emulated else torch.cuda.Stream(device=self.grads.device)
        if not self.emulated:
            # the per-bucket all-reduces are ordered against torch's current stream: the library must enqueue there too
            engine.set_stream(torch.cuda.current_stream(self.grads.device).cuda_stream)
        engine.set_grad_buckets(True)

    def reduce_async(self, ntok_t):
        """Enqueue the per-bucket all-reduces (side stream, each waiting on its bucket's event) and the all-reduce of the
        1-element float32 device tensor ntok_t (this rank's unmasked-token count, summed in place); the current stream then
        waits for both.  Nothing synchronises with the host: follow with engine.adam_ema_step_dev(ntok_t)."""
        import torch
        import torch.distributed as dist
        buckets = self.engine.grad_buckets()
        if self.emulated:
            for off, n in buckets:
                dist.all_reduce(self.grads[off:off + n])
            dist.all_reduce(ntok_t)
            return
        main = torch.cuda.current_stream(self.grads.device)
        for i, (off, n) in enumerate(buckets):
            self.engine.grad_bucket_wait(i, self.comm.cuda_stream)      # device-side wait on the bucket's event
            with torch.cuda.stream(self.comm):
                dist.all_reduce(self.grads[off:off + n])
        main.wait_stream(self.comm)
        dist.all_reduce(ntok_t)

    def reduce(self, ntok_t):
        """reduce_async + read the global token count back (host synchronisation)."""
        self.reduce_async(ntok_t)
        return float(ntok_t.item())
