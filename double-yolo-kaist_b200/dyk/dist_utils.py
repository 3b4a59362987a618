"""Multi-GPU plumbing (torch.distributed; NCCL on the GPU box, gloo in the CPU tests).

The path shards by frame: every rank holds a full replica and its own slice of the batch.  Inference and NMS need no
data-path collective at all.  Training adds exactly one exchange per optimizer step — the gradient all-reduce the
reference gets from DistributedDataParallel (train.py:124-126) — done here over a few large flat buckets so that NCCL
runs at NVLink bandwidth instead of paying one launch per parameter tensor (568 of them in dyolov4_fshare).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_ranks():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend: str = None, device=None) -> int:
    """Initialises the default process group from the torchrun environment (no-op for world size 1)."""
    rank, local, world = env_ranks()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return world


def barrier(device=None) -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(value: float, device="cpu") -> float:
    """A device-timed duration is only meaningful as the maximum over ranks."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_seed(rank: int, slot: int) -> int:
    """Distinct synthetic frames per rank and per ring slot."""
    return 1000 * rank + slot


def frame_shard(n_frames: int, rank: int, world: int):
    """[begin, end) of the frames rank `rank` owns when n_frames are split as evenly as possible."""
    base, rem = divmod(n_frames, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class OverlappedAllReduce:
    """Gradient averaging that overlaps the backward pass.

    `model.grad_reducer = OverlappedAllReduce()` makes the training plan (dyk/train_plan.py) report, while it enqueues
    the backward launches, which ranges of its flat fp32 gradient buffer are final; every time ~bucket_bytes of
    contiguous finished gradients have accumulated an asynchronous all-reduce of exactly that slice is issued (NCCL runs
    it on its own stream behind the kernels enqueued so far), and the end of backward waits for all of them on the
    stream, not on the host.  The parameter gradients autograd hands out afterwards are views of the reduced buffer, so
    neither DistributedDataParallel nor allreduce_gradients() is needed (and must not be used on top).
    Backward finishes parameters from the last layer to the first, i.e. from the end of the flat buffer downwards."""

    def __init__(self, bucket_bytes: int = 48 << 20, group=None, comm_dtype=None):
        """comm_dtype=torch.bfloat16 sends every bucket as bf16 (half the NVLink bytes; the sum is then formed in bf16 by
        NCCL, i.e. with 8 mantissa bits — an opt-in trade, the default keeps the reference's fp32 gradient exchange)."""
        self.bucket_bytes, self.group, self.comm_dtype = bucket_bytes, group, comm_dtype
        self.calls = 0
        self.steps = 0

    def _active(self):
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _native_avg(self):
        return self._active() and dist.get_backend(self.group) == "nccl"      # gloo has no ReduceOp.AVG

    def begin(self, flat: torch.Tensor, numel: int) -> None:
        self.flat, self.numel = flat, numel
        self.ready = []            # finished (lo, hi) ranges not yet sent
        self.sent_lo = numel       # everything in [sent_lo, numel) has been handed to a collective
        self.works = []
        self.avg = self._native_avg()
        self.steps += 1

    def feed(self, ranges) -> None:
        if not self._active() or not ranges:
            return
        self.ready.extend(ranges)
        self._flush(final=False)

    def _flush(self, final: bool) -> None:
        # grow the contiguous finished region that ends at sent_lo; send it when it is big enough (or at the end)
        self.ready.sort()
        lo = self.sent_lo
        while self.ready and self.ready[-1][1] >= lo:
            lo = min(lo, self.ready.pop()[0])
        if lo < self.sent_lo and (final or (self.sent_lo - lo) * 4 >= self.bucket_bytes):
            self._send(lo, self.sent_lo)
            self.sent_lo = lo
        elif lo < self.sent_lo:
            self.ready.append((lo, self.sent_lo))     # keep the merged region for the next call

    def _send(self, lo: int, hi: int) -> None:
        chunk = self.flat[lo:hi]
        op = dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM
        wire = chunk if self.comm_dtype is None else chunk.to(self.comm_dtype)
        self.works.append((dist.all_reduce(wire, op=op, group=self.group, async_op=True), chunk, wire))
        self.calls += 1

    def finish(self) -> None:
        if not self._active():
            return
        self._flush(final=True)
        if self.ready:             # ranges that never became contiguous with the sent region (should not happen)
            for lo, hi in self.ready:
                self._send(lo, hi)
            self.ready = []
        world = dist.get_world_size(self.group)
        for work, chunk, wire in self.works:
            work.wait()
            if wire is not chunk:
                chunk.copy_(wire)
            if not self.avg:
                chunk.div_(world)
        self.works = []


def _flat_view(chunk):
    """The one contiguous tensor behind `chunk` when its tensors are back-to-back views of one storage (what
    dyk.train_plan hands out, padded to 32 floats), else None."""
    first = chunk[0]
    if any(g.untyped_storage().data_ptr() != first.untyped_storage().data_ptr() or not g.is_contiguous() for g in chunk):
        return None
    lo = first.storage_offset()
    hi = chunk[-1].storage_offset() + chunk[-1].numel()
    if any(a.storage_offset() > b.storage_offset() for a, b in zip(chunk, chunk[1:])):
        return None
    if hi - lo > sum(g.numel() for g in chunk) + 32 * len(chunk):
        return None
    base = torch.empty(0, dtype=first.dtype, device=first.device).set_(first.untyped_storage(), lo, (hi - lo,), (1,))
    return base


def allreduce_gradients(params, bucket_bytes: int = 64 << 20) -> int:
    """Averages .grad of `params` over all ranks through flat buckets of ~bucket_bytes.  Returns the number of
    collectives issued.  Gradients that are consecutive views of one flat buffer (what dyk.train_plan produces) are
    reduced in place without any copy; others are gathered into a temporary bucket."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return 0
    world = dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    calls, i = 0, 0
    while i < len(grads):
        j, nbytes = i, 0
        while j < len(grads) and (j == i or nbytes + grads[j].numel() * grads[j].element_size() <= bucket_bytes):
            if grads[j].dtype != grads[i].dtype or grads[j].device != grads[i].device:
                break
            nbytes += grads[j].numel() * grads[j].element_size()
            j += 1
        chunk = grads[i:j]
        inplace = _flat_view(chunk) if len(chunk) > 1 else None
        if inplace is not None:
            dist.all_reduce(inplace, op=dist.ReduceOp.SUM)
            inplace.div_(world)
            calls += 1
            i = j
            continue
        flat = torch.cat([g.reshape(-1) for g in chunk]) if len(chunk) > 1 else chunk[0].reshape(-1)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        if len(chunk) > 1 or flat.data_ptr() != chunk[0].data_ptr():     # reshape(-1) of a non-contiguous grad is a copy
            off = 0
            for g in chunk:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        calls += 1
        i = j
    return calls
