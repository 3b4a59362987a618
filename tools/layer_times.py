#!/usr/bin/env python
"""Per-launch CUDA-event timing of a model's plan (eager replay), one line per native step.
    python tools/layer_times.py [cfg] [batch] [out.json]"""
import json, sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
import bench
import models
from dyk.plan import _ConvStep

cfg = sys.argv[1] if len(sys.argv) > 1 else "kaist_dyolov3_add_sl.cfg"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/layers.json"
path, ref, st = bench.oracle_objects(cfg)
m = models.YOLO(path, (512, 640)); m.load_state_dict(st); m = m.cuda().eval()
v, l = [t.cuda() for t in bench.synthetic_frames(B, 0)]
dual = "second_index" in m.net_info
with torch.no_grad():
    m(v, l) if dual else m(v)
plan = m._plans.last_plan
rows = []
iters = 5
acc = [0.0] * len(plan.steps)
for it in range(iters + 1):
    plan.run_stems(v, l if dual else None)
    evs = []
    for s in plan.steps:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); s(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    if it:
        for i, (a, b) in enumerate(evs):
            acc[i] += a.elapsed_time(b) / iters
for s, ms in zip(plan.steps, acc):
    if isinstance(s, _ConvStep):
        k = s.kw["k"]; up = 4 if s.kw["upsample2x"] else 1
        opix = s.y.N * s.y.H * s.y.W // up
        cout = s.e["conv"].out_channels
        fl = 2.0 * opix * cout * s.x.C * k * k
        by = 2.0 * (s.x.N * s.x.H * s.x.W * s.x.C + s.y.N * s.y.H * s.y.W * cout + cout * s.x.C * k * k
                    + (opix * cout if s.kw["res"] is not None else 0))
        rows.append(dict(kind="conv", cin=s.x.C, cout=cout, k=k, stride=s.kw["stride"], H=s.x.H, W=s.x.W, res=s.kw["res"] is not None,
                         up=s.kw["upsample2x"], ms=ms, tflops=fl / ms / 1e9, gbs=by / ms / 1e6, gflop=fl / 1e9))
    elif hasattr(s, "fn"):
        rows.append(dict(kind=getattr(s.fn, "__name__", "?"), ms=ms))
Path(out).parent.mkdir(exist_ok=True)
json.dump(rows, open(out, "w"), indent=0)
agg = {}
for r in rows:
    key = (r["kind"], r.get("cin"), r.get("cout"), r.get("k"), r.get("stride"), r.get("H"), r.get("res"), r.get("up"))
    a = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += r["ms"]; a[2] += r.get("gflop", 0.0); a[3] = r.get("gbs", 0.0)
tot = sum(r["ms"] for r in rows)
print(f"total {tot:.3f} ms over {len(rows)} launches")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf = a[2] / a[1] if a[1] else 0
    print(f"{a[1]:8.3f} ms {100 * a[1] / tot:5.1f}%  n={a[0]:3d}  {tf:7.1f} TF/s {a[3]:7.0f} GB/s  {key}")
