#!/usr/bin/env python
"""A few launches of the tensor-core stem (3 -> 32, 3x3, 512x640 x 16, uint8 frames) for `ncu --set full -k regex:stem_tc`."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
from dyk import ops
g = torch.Generator().manual_seed(0)
x = torch.randint(0, 256, (16, 3, 512, 640), dtype=torch.uint8, generator=g).cuda()
w = (torch.randn((32, 3, 3, 3), generator=g) / 27 ** 0.5).cuda().permute(0, 2, 3, 1).contiguous()
sc, bi = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
y = ops.new_view(16, 512, 640, 32, torch.float16, torch.device("cuda"))
for _ in range(3):
    ops.nhwc_stem(x, w, sc, bi, y, k=3, stride=1, pad=1, act="leaky")
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.nhwc_stem(x, w, sc, bi, y, k=3, stride=1, pad=1, act="leaky")
b.record()
torch.cuda.synchronize()
print(f"stem 3->32 @512x640 x16 uint8: {a.elapsed_time(b) / 10 * 1e3:.1f} us per launch (HBM floor 87 us)")
