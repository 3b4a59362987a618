#!/usr/bin/env python
"""A few launches of the two MobileNet-path kernels at MobileNetV3-dual bs-64 shapes for `ncu --set full`:
the TMA-staged depthwise kernel (3x3 C=64 @128x160, 5x5 C=120 @64x80) and the thin 1x1 warp-MMA kernel
(24->72 @128x160, 16->16 @256x320), the unrolled stride-2 stem (uint8 3->16 @512x640) and the SPP pool (512 ch @16x20), fp16."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
from dyk import ops
from dyk.ops import View
dt, N = torch.float16, 64


def dw(C, H, W, k, s, act):
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    x = View(torch.randn((N, H, W, C), device="cuda").to(dt), 0, C)
    y = View(torch.empty((N, Ho, Wo, C), device="cuda", dtype=dt), 0, C)
    w = torch.randn((k, k, C), device="cuda")
    sc, bi = torch.ones(1024, device="cuda"), torch.zeros(1024, device="cuda")
    for _ in range(2):
        ops.nhwc_dwconv(x, w, sc, bi, y, k=k, stride=s, pad=k // 2, act=act)


def thin(Cin, H, W, Cout, act):
    x = View(torch.randn((N, H, W, Cin), device="cuda").to(dt), 0, Cin)
    y = View(torch.empty((N, H, W, Cout), device="cuda", dtype=dt), 0, Cout)
    w = (torch.randn((Cout, 1, 1, Cin), device="cuda") / Cin ** 0.5).to(dt)
    sc, bi = torch.ones(1024, device="cuda"), torch.zeros(1024, device="cuda")
    for _ in range(2):
        ops.nhwc_conv(x, w, sc, bi, y, k=1, stride=1, pad=0, act=act)


def stem_and_pool():
    x = torch.randint(0, 256, (N, 3, 512, 640), dtype=torch.uint8, device="cuda")
    y = View(torch.empty((N, 256, 320, 16), device="cuda", dtype=dt), 0, 16)
    w = torch.randn((16, 3, 3, 3), device="cuda") * 0.3
    sc, bi = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")
    for _ in range(2):
        ops.nhwc_stem(x, w, sc, bi, y, k=3, stride=2, pad=1, act="hard-swish")
    p = View(torch.randn((N, 16, 20, 512), device="cuda").to(dt), 0, 512)
    q = View(torch.empty((N, 16, 20, 512), device="cuda", dtype=dt), 0, 512)
    for _ in range(2):
        ops.nhwc_maxpool(p, q, 5, 1)


dw(64, 128, 160, 3, 1, "relu")
dw(120, 64, 80, 5, 1, "relu")
thin(24, 128, 160, 72, "relu")
thin(16, 256, 320, 16, "relu")
stem_and_pool()
torch.cuda.synchronize()
print("done")
