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


def allreduce_gradients(params, bucket_bytes: int = 64 << 20) -> int:
    """Averages .grad of `params` over all ranks through flat buckets of ~bucket_bytes.  Returns the number of
    collectives issued.  Gradients that are consecutive views of one flat buffer (what dyk.train_plan produces) are
    reduced in place without any copy."""
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
        flat = torch.cat([g.reshape(-1) for g in chunk]) if len(chunk) > 1 else chunk[0].reshape(-1)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        if len(chunk) > 1:
            off = 0
            for g in chunk:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        calls += 1
        i = j
    return calls
