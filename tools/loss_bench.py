#!/usr/bin/env python
"""Native compute_loss (forward + backward) against the oracle port of the reference loss run in eager PyTorch on the same
GPU, at BASELINE size: bs16, 512x640 heads, 3 labels per frame.
    python tools/loss_bench.py"""
import sys, time, types
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
from build_utils.utils import compute_loss
from dyk import _native as nat
from oracle import loss_ref

dev = "cuda:0"
g = torch.Generator().manual_seed(1)
B, nt = 16, 48
anchors_px = torch.tensor([[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]], dtype=torch.float32)
strides, masks = [8, 16, 32], [[0, 1, 2], [3, 4, 5], [6, 7, 8]]
anchors = [anchors_px[m] / s for m, s in zip(masks, strides)]
p = [torch.randn((B, 3, 512 // s, 640 // s, 6), generator=g).to(dev).requires_grad_(True) for s in strides]
t = torch.zeros((nt, 6))
t[:, 0] = torch.arange(nt) // 3
t[:, 2:4] = torch.rand((nt, 2), generator=g) * 0.8 + 0.1
t[:, 4] = torch.rand((nt,), generator=g) * 0.07 + 0.03
t[:, 5] = torch.rand((nt,), generator=g) * 0.2 + 0.1
hyp = {"box": 3.54, "cls": 37.4, "obj": 64.3, "cls_pw": 1.0, "obj_pw": 1.0, "iou_t": 0.20, "fl_gamma": 0.0, "ciou": 1.0}
model = types.SimpleNamespace(hyp=hyp, gr=1.0, nc=1, cfg="kaist_dyolov4", yolo_layers=[0, 1, 2],
                              module_list=[types.SimpleNamespace(anchor_vec=a) for a in anchors])
td = t.to(dev)


def native():
    parts = compute_loss(p, td, model)
    (parts["box_loss"] + parts["obj_loss"] + parts["class_loss"]).backward()


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


l0 = nat.launch_count()
native()
ms = timeit(native)
l0 = nat.launch_count()
native()
print(f"native compute_loss forward+backward: {ms:.3f} ms per step ({nat.launch_count() - l0} native launches per step)")
# oracle port in eager torch on the GPU (the reference's own function mixes CPU index tensors with CUDA tensors, utils.py:336,361)
torch.set_default_device(dev)
pa = [a.to(dev) for a in anchors]


def eager():
    lb, lo, lc = loss_ref.compute_loss(p, td, pa, hyp, 1.0, 1, True)
    (lb + lo + lc).backward()


try:
    print(f"oracle port of the reference loss, eager PyTorch on the same GPU: {timeit(eager, 10):.3f} ms per step")
except Exception as e:  # noqa: BLE001
    print("eager oracle port on GPU failed:", type(e).__name__, e)
