#!/usr/bin/env python
"""A few launches of the resident-weights pair kernel (32->64 3x3 @256x320, bs 16) for `ncu --set full`."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
from dyk import ops
from dyk.ops import View
dt = torch.float16
N, Cin, H, W, Cout = 16, 32, 256, 320, 64
x = View(torch.randn((N, H, W, Cin), device="cuda").to(dt), 0, Cin)
y = View(torch.empty((N, H, W, Cout), device="cuda", dtype=dt), 0, Cout)
w = (torch.randn((Cout, 3, 3, Cin), device="cuda") / (Cin * 9) ** 0.5).to(dt)
sc, bi = torch.ones(2048, device="cuda"), torch.zeros(2048, device="cuda")
for _ in range(3):
    ops.nhwc_conv(x, w, sc, bi, y, k=3, stride=1, pad=1, act="leaky", res=None)
torch.cuda.synchronize()
print("done")
