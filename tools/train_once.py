#!/usr/bin/env python
"""One eager training step of a model at BASELINE size — the command ncu wraps for the training launch list.
    python tools/train_once.py [cfg] [batch] [steps]"""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
import bench, models
from dyk import cfg_zoo
cfg = sys.argv[1] if len(sys.argv) > 1 else "kaist_dyolov4_fshare_global_concat_se3.cfg"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
m = models.YOLO(cfg_zoo.materialize(cfg), (512, 640)).cuda().train()
m.compute_dtype = torch.bfloat16
v, l = [t.cuda() for t in bench.synthetic_frames(B, 0)]
dual = "second_index" in m.net_info
for _ in range(steps):
    p = m(v, l) if dual else m(v)
    sum((t.float() ** 2).mean() for t in p).backward()
torch.cuda.synchronize()
print("done")
