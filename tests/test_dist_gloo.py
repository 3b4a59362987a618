"""World-size-2 checks of the multi-GPU plumbing on CPU (gloo): sharding, max-over-ranks timing, gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys
    from pathlib import Path
    repo = Path(__file__).resolve().parent.parent
    sys.path[:0] = [str(repo), str(repo / "double-yolo-kaist_b200")]
    from dyk import dist_utils as du
    assert du.init("gloo") == world
    # timing: max over ranks
    assert du.max_over_ranks(10.0 + rank) == 10.0 + world - 1
    # frames shard without gaps or overlap
    spans = [du.frame_shard(37, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == 37 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert du.shard_seed(rank, 3) != du.shard_seed((rank + 1) % world, 3)
    # gradient all-reduce: several parameters, small buckets so that more than one collective is used
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 300, 7, 1024)]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    calls = du.allreduce_gradients(params, bucket_bytes=2048)
    assert calls >= 2
    mean = sum(range(1, world + 1)) / world
    for i, p in enumerate(params):
        assert torch.allclose(p.grad, torch.full_like(p, mean * (i + 1)))
    # contiguous views of one flat buffer are reduced in place (no temporary bucket)
    flat = torch.arange(96, dtype=torch.float32) * (rank + 1)
    views = [torch.nn.Parameter(torch.zeros(n)) for n in (32, 20, 32)]
    for p_, (o, n) in zip(views, ((0, 32), (32, 20), (64, 32))):
        p_.grad = flat[o:o + n]
    assert du.allreduce_gradients(views, bucket_bytes=1 << 20) == 1
    assert torch.allclose(flat[:52], torch.arange(52, dtype=torch.float32) * mean)
    # overlapped all-reduce: ranges become final from the end of the buffer downwards (and slightly out of order);
    # buckets are sent as soon as enough contiguous finished gradients exist, the rest at finish()
    n = 4096
    flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
    red = du.OverlappedAllReduce(bucket_bytes=1024 * 4)
    red.begin(flat, n)
    red.feed([(3584, 4096)])
    red.feed([(2048, 3072)])                 # not contiguous with the sent region yet
    red.feed([(3072, 3584)])                 # closes the gap: [2048, 3584) can go
    assert red.calls >= 1 and red.sent_lo == 2048
    red.feed([(0, 1000)])
    red.feed([(1000, 2048)])
    red.finish()
    assert red.sent_lo == 0 and not red.ready
    assert torch.allclose(flat, torch.arange(n, dtype=torch.float32) * mean)
    du.barrier()
    out.put(rank)


def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(out.get(timeout=5) for _ in range(world)) == list(range(world))
