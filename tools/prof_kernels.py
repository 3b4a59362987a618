#!/usr/bin/env python
"""A handful of launches of the dominant kernels at BASELINE shapes for `ncu --set full`:
the CTA-pair halo conv (256->512 3x3 @32x40), the generic conv (256->128 1x1 @64x80, 32->64 3x3 s2 @512x640) and the
tcgen05 wgrad (256->512 3x3 @32x40), batch 16, fp16."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
from dyk import ops, train_ops as T
from dyk.ops import View
dt = torch.float16


def conv(N, Cin, H, W, Cout, k, s, res):
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    x = View(torch.randn((N, H, W, Cin), device="cuda").to(dt), 0, Cin)
    y = View(torch.empty((N, Ho, Wo, Cout), device="cuda", dtype=dt), 0, Cout)
    r = View(torch.randn((N, Ho, Wo, Cout), device="cuda").to(dt), 0, Cout) if res else None
    w = (torch.randn((Cout, k, k, Cin), device="cuda") / (Cin * k * k) ** 0.5).to(dt)
    sc, bi = torch.ones(2048, device="cuda"), torch.zeros(2048, device="cuda")
    for _ in range(3):
        ops.nhwc_conv(x, w, sc, bi, y, k=k, stride=s, pad=pad, act="leaky", res=r)
    return x, y


x, y = conv(16, 256, 32, 40, 512, 3, 1, True)
conv(16, 256, 64, 80, 128, 1, 1, False)
conv(16, 32, 512, 640, 64, 3, 2, False)
grad = torch.empty((512, 256, 3, 3), device="cuda")
for _ in range(3):
    T.conv_wgrad(x, y, grad, k=3, stride=1, pad=1, accumulate=False)
torch.cuda.synchronize()
print("done")
