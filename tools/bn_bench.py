#!/usr/bin/env python
"""Achieved HBM bandwidth of the train-mode BatchNorm kernels on the shapes that dominate the dyolov4 training step.
    python tools/bn_bench.py [act]
Each call is timed with CUDA events over `iters` repetitions on buffers larger than the 126 MB L2 where the real layer is."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
from dyk import ops, train_ops as T
from dyk.ops import View

act = sys.argv[1] if len(sys.argv) > 1 else "mish"
DEV = "cuda:0"
dt = torch.bfloat16
shapes = [(16, 512, 640, 32), (16, 256, 320, 64), (16, 256, 320, 32), (16, 128, 160, 128), (16, 128, 160, 64), (16, 64, 80, 256)]


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3      # us


print(f"{'shape':>22s} {'MB':>6s} | {'stats 1E':>14s} | {'apply 2E':>14s} | {'bwd 5-6E':>14s} | {'axpby 2E':>14s}")
for N, H, W, C in shapes:
    z = View(torch.randn(N, H, W, C, device=DEV).to(dt), 0, C)
    y = ops.new_view(N, H, W, C, dt, DEV)
    dy = View(torch.randn(N, H, W, C, device=DEV).to(dt), 0, C)
    dz = ops.new_view(N, H, W, C, dt, DEV)
    g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    sc, sh, mu, inv = (torch.empty(C, device=DEV) for _ in range(4))
    dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    E = z.buf.numel() * 2 / 1e6
    t1 = timed(lambda: T.bn_train_stats(z, g, b, 1e-5, 0.1, rm, rv, sc, sh, mu, inv))
    t2 = timed(lambda: T.bn_act_apply(z, sc, sh, act, y))
    t3 = timed(lambda: T.bn_act_bwd(dy, z, sc, sh, mu, inv, g, act, dz, dg, db))
    t4 = timed(lambda: T.axpby(dy, dz, None, False))
    nb = 6 if act == "mish" else 5
    f = lambda t, n: f"{t:6.1f}us {n * E / t:4.2f}TB/s"
    print(f"{str((N, H, W, C)):>22s} {E:6.0f} | {f(t1, 1)} | {f(t2, 2)} | {f(t3, nb)} | {f(t4, 2)}")
