#!/usr/bin/env python
"""One eager (non-graph) forward of a model at BASELINE size — the command ncu wraps for launch lists."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
import bench, models
from build_utils.utils import nms_raw
cfg = sys.argv[1] if len(sys.argv) > 1 else "kaist_dyolov3_add_sl.cfg"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
path, ref, st = bench.oracle_objects(cfg)
m = models.YOLO(path, (512, 640)); m.load_state_dict(st); m = m.cuda().eval()
m.use_cuda_graph = False
v, l = [t.cuda() for t in bench.synthetic_frames(B, 0)]
dual = "second_index" in m.net_info
for _ in range(reps):
    with torch.no_grad():
        io, _ = m(v, l) if dual else m(v)
    nms_raw(io, 0.01, 0.6, False, None, False, 100)
torch.cuda.synchronize()
print("done")
