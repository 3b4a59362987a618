#!/usr/bin/env python
"""Micro-benchmark of the depthwise convolution kernel on the MobileNetV3-dual layer shapes (bs 64, fp16; CUDA events, L2
flushed between runs).    python tools/dw_bench.py [--only i,j]"""
import argparse, os, sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
from dyk import ops
from dyk.ops import View
SHAPES = [(72, 128, 160, 3, 1, "relu"), (16, 256, 320, 3, 1, "relu"), (64, 128, 160, 3, 1, "relu"), (120, 64, 80, 5, 1, "relu"),
          (128, 64, 80, 3, 1, "relu6"), (672, 32, 40, 3, 1, "hard-swish"), (256, 32, 40, 3, 1, "relu6"), (480, 32, 40, 3, 1, "hard-swish"),
          (960, 16, 20, 5, 1, "hard-swish"), (128, 64, 80, 3, 2, "relu6"), (672, 16, 20, 5, 1, "hard-swish")]
ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--modes", default="1,0")
ap.add_argument("--dump", default="")
ap.add_argument("--sweep", action="store_true", help="time every tile geometry (DYK_DW_TILE_CFG) and list the best next to the model's pick")
args = ap.parse_args()
sel = [int(i) for i in args.only.split(",")] if args.only else range(len(SHAPES))
N, dt = 64, torch.float16
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for i in sel:
    C, H, W, k, s, act = SHAPES[i]
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    x = View(torch.randn((N, H, W, C), device="cuda").to(dt), 0, C)
    y = View(torch.empty((N, Ho, Wo, C), device="cuda", dtype=dt), 0, C)
    w = torch.randn((k, k, C), device="cuda")
    sc, bi = torch.ones(1024, device="cuda"), torch.zeros(1024, device="cuda")
    def timed(iters):
        ts = []
        for it in range(iters + 2):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.nhwc_dwconv(x, w, sc, bi, y, k=k, stride=s, pad=k // 2, act=act); b.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(a.elapsed_time(b) * 1e3)
        return sorted(ts)[len(ts) // 2]
    if args.sweep:
        os.environ["DYK_DW_TILE"] = "1"
        os.environ.pop("DYK_DW_TILE_CFG", None)
        pick = timed(args.iters)
        rows = []
        for CB in [c for c in range(8, 257, 8) if C % c == 0 and (c >= 32 or c == C)] + ([64] if C > 256 else []):
            for TW in (8, 12, 16, 20, 24, 28, 32, 40, 48, 64):
                if TW >= Wo + 4:
                    continue
                for TH in range(1, min(16, Ho) + 1):
                    os.environ["DYK_DW_TILE_CFG"] = f"{CB},{TW},{TH}"
                    try:
                        rows.append((timed(3), CB, TW, TH))
                    except Exception:
                        break
        os.environ.pop("DYK_DW_TILE_CFG", None)
        rows.sort()
        if args.dump:
            import json
            with open(args.dump, "a") as f:
                f.write(json.dumps({"shape": [C, H, W, k, s, N], "pick_us": pick, "rows": rows}) + "\n")
        by = 2.0 * N * (H * W * C + Ho * Wo * C)
        print(f"[{i:2d}] C={C:4d} {H}x{W} k{k} s{s}: model pick {pick:7.1f} us {by / pick / 1e3:5.0f} GB/s; best of {len(rows)}: "
              + "  ".join(f"{us:6.1f} us (CB {cb} TW {tw} TH {th})" for us, cb, tw, th in rows[:6]), flush=True)
        continue
    res = []
    for mode in args.modes.split(","):              # 1 = TMA-staged tile kernel, 0 = strip kernel (DYK_DW_TILE)
        os.environ["DYK_DW_TILE"] = mode
        ts = []
        for it in range(args.iters + 2):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.nhwc_dwconv(x, w, sc, bi, y, k=k, stride=s, pad=k // 2, act=act); b.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(a.elapsed_time(b) * 1e3)
        res.append(sorted(ts)[len(ts) // 2])
    by = 2.0 * N * (H * W * C + Ho * Wo * C)
    cols = "  ".join(f"{'tile' if m == '1' else 'strip'} {us:7.1f} us {by / us / 1e3:5.0f} GB/s" for m, us in zip(args.modes.split(","), res))
    print(f"[{i:2d}] C={C:4d} {H}x{W} k{k} s{s} {act:10s}: {cols}  (HBM floor {by / 6.4e12 * 1e6:5.1f} us)", flush=True)
