#!/usr/bin/env python
"""Micro-benchmark of dyk_conv2d_fwd on the BASELINE layer shapes (CUDA events, L2 flushed between runs).
With DYK_B200_LIB=double-yolo-kaist_b200/libdyk_b200_prof.so (python double-yolo-kaist_b200/build.py --profile)
it also prints where the producer / MMA / epilogue roles of the kernel wait.
    python tools/conv_bench.py [--only i,j] [--iters n] [--dtype fp16|bf16]"""
import argparse, sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
import ctypes as C
import os
PROFILE = os.environ.get("DYK_B200_LIB", "").endswith("_prof.so")   # role-cycle counters need the profile build
from dyk import ops
from dyk import _native as nat
from dyk.ops import View

SHAPES = [  # (N, Cin, H, W, Cout, k, stride, res, act)
    (16, 256, 64, 80, 128, 1, 1, False, "leaky"),
    (16, 128, 64, 80, 256, 3, 1, True, "leaky"),
    (16, 256, 32, 40, 512, 3, 1, True, "leaky"),
    (16, 512, 16, 20, 1024, 3, 1, True, "leaky"),
    (16, 64, 128, 160, 128, 3, 1, True, "leaky"),
    (16, 32, 256, 320, 64, 3, 1, True, "leaky"),
    (16, 32, 512, 640, 64, 3, 2, False, "leaky"),
    (16, 512, 32, 40, 256, 1, 1, False, "leaky"),
    (16, 128, 128, 160, 64, 1, 1, False, "leaky"),
    (16, 64, 256, 320, 32, 1, 1, False, "leaky"),
    (16, 1024, 16, 20, 512, 1, 1, False, "leaky"),
    (16, 512, 32, 40, 512, 3, 1, False, "leaky"),
    (16, 128, 64, 80, 128, 3, 1, True, "mish"),
    (16, 64, 256, 320, 64, 1, 1, False, "mish"),
]

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--dtype", default="fp16")
ap.add_argument("--noflush", action="store_true")
args = ap.parse_args()
dt = torch.float16 if args.dtype == "fp16" else torch.bfloat16
sel = [int(i) for i in args.only.split(",")] if args.only else range(len(SHAPES))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for i in sel:
    N, Cin, H, W, Cout, k, s, res, act = SHAPES[i]
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    x = View(torch.randn((N, H, W, Cin), device="cuda").to(dt), 0, Cin)
    y = View(torch.empty((N, Ho, Wo, Cout), device="cuda", dtype=dt), 0, Cout)
    r = View(torch.randn((N, Ho, Wo, Cout), device="cuda").to(dt), 0, Cout) if res else None
    w = (torch.randn((Cout, k, k, Cin), device="cuda") / (Cin * k * k) ** 0.5).to(dt)
    scale = torch.ones(1024 if Cout <= 1024 else 2048, device="cuda")
    bias = torch.zeros_like(scale)
    times = []
    for it in range(args.iters + 2):
        if not args.noflush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.nhwc_conv(x, w, scale, bias, y, k=k, stride=s, pad=pad, act=act, res=r)
        b.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(a.elapsed_time(b) * 1e3)
    us = sorted(times)[len(times) // 2]
    fl = 2.0 * N * Ho * Wo * Cout * Cin * k * k
    by = 2.0 * (N * H * W * Cin + N * Ho * Wo * Cout * (2 if res else 1) + Cout * Cin * k * k)
    role = ""
    if PROFILE:
        prof = torch.zeros(16, dtype=torch.int64, device="cuda"); prof[8] = 1 << 62; prof[12] = 1 << 62   # 8..13: globaltimer stamps
        nat.call("dyk_conv_set_profile", C.c_void_p(prof.data_ptr()))
        ops.nhwc_conv(x, w, scale, bias, y, k=k, stride=s, pad=pad, act=act, res=r)
        torch.cuda.synchronize()
        nat.call("dyk_conv_set_profile", None)
        pr = prof.tolist()
        n = max(pr[7], 1)
        role = (f"prod wait-empty {pr[0] / max(pr[1], 1):4.0%} | mma wait-data {pr[2] / max(pr[4], 1):4.0%} wait-acc "
                f"{pr[3] / max(pr[4], 1):4.0%} | epi wait-acc {pr[5] / max(pr[6], 1):4.0%} | cyc/CTA {pr[4] / n:8.0f}")
    print(f"[{i:2d}] {Cin:5d}->{Cout:5d} k{k} s{s} {H}x{W} res={int(res)} {act:6s}: {us:8.1f} us  {fl / us / 1e6:7.1f} TF/s  {by / us / 1e3:7.0f} GB/s"
          f"   (roofline {max(fl / 1.59e15, by / 6.65e12) * 1e6:6.1f} us)  {role}", flush=True)
