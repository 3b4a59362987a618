#!/usr/bin/env python
"""Per-step CUDA-event timing of one training iteration (forward steps and backward steps of the TrainPlan).
    python tools/train_times.py [cfg] [batch] [H] [W]"""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
import bench, models
from dyk import cfg_zoo, plan as P

cfg = sys.argv[1] if len(sys.argv) > 1 else "kaist_dyolov4_fshare_global_concat_se3.cfg"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
H = int(sys.argv[3]) if len(sys.argv) > 3 else 512
W = int(sys.argv[4]) if len(sys.argv) > 4 else 640
m = models.YOLO(cfg_zoo.materialize(cfg), (H, W)).cuda().train()
m.compute_dtype = torch.bfloat16
v, l = [t.cuda() for t in bench.synthetic_frames(B, 0)] if (H, W) == (512, 640) else [torch.rand(B, 3, H, W).cuda() for _ in range(2)]
dual = "second_index" in m.net_info
for _ in range(2):
    p = m(v, l) if dual else m(v)
    sum((t ** 2).mean() for t in p).backward()
plan = m._train_plans.last_plan
torch.cuda.synchronize()


def describe(f):
    d = f.__defaults__ or ()
    c = getattr(f, "__closure__", None) or ()
    objs = list(d) + [x.cell_contents for x in c if True]
    for o in objs:
        if isinstance(o, dict) and "op" in o:
            op, conv = o["op"], o["conv"]
            return f"conv L{op.layer} {conv.in_channels}->{conv.out_channels} k{o['k']} s{o['s']} {op.out.H}x{op.out.W}"
        if isinstance(o, P.Op):
            return f"{o.kind} L{o.layer}"
    return getattr(f, "__name__", "?")


def timed(fns, call):
    evs = []
    for f in fns:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(f); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in evs]


ft = timed(plan.fwd, lambda f: f(v, l if dual else None))
flat = torch.zeros(plan.grad_numel, device="cuda")
dps = [torch.randn_like(t) for t in plan.p_outs]
bt = timed(plan.bwd, lambda f: f(flat, dps))
print(f"forward {sum(ft):.2f} ms in {len(ft)} steps, backward {sum(bt):.2f} ms in {len(bt)} steps")


def conv_info(f):
    """(pixels_out, Cin, Cout, k, stride) of the convolution block behind a forward / backward step, or None"""
    for o in list(f.__defaults__ or ()) + [c.cell_contents for c in (getattr(f, "__closure__", None) or ())]:
        if isinstance(o, dict) and "op" in o:
            op, conv = o["op"], o["conv"]
            return dict(layer=op.layer, cin=conv.in_channels, cout=conv.out_channels, k=o["k"], s=o["s"], groups=conv.groups,
                        Ho=op.out.H, Wo=op.out.W, Hi=op.src.H, Wi=op.src.W, act=op.act)
    return None


import json, os
rows = []
for name, ts, fns in (("fwd", ft, plan.fwd), ("bwd", bt, plan.bwd)):
    for t, f in zip(ts, fns):
        ci = conv_info(f)
        row = dict(phase=name, ms=t, what=describe(f))
        if ci:
            eo = B * ci["Ho"] * ci["Wo"] * ci["cout"] * 2 / 1e9          # GB of one 16-bit pass over the output
            ei = B * ci["Hi"] * ci["Wi"] * ci["cin"] * 2 / 1e9
            gf = 2.0 * B * ci["Ho"] * ci["Wo"] * ci["cout"] * ci["cin"] // ci["groups"] * ci["k"] ** 2 / 1e9
            if name == "fwd":     # conv (in + z) + stats (z) + apply (z + y)
                gb, fl = ei + 4 * eo, gf
            else:                 # BN bwd reduce (dy, z [+ g]) + apply (g, z, dz) + wgrad (x, dz) + dgrad (dz, dx)
                gb, fl = (6 if ci["act"] == "mish" else 5) * eo + (ei + eo) + (eo + ei), 2 * gf
            row.update(ci, gb=gb, gflop=fl, ideal_ms=max(gb / 6.4, fl / 1400.0))
        rows.append(row)
out = os.environ.get("DYK_TRAIN_TIMES_JSON")
if out:
    json.dump(rows, open(out, "w"))
tot = sum(r["ms"] for r in rows if "ideal_ms" in r)
ideal = sum(r["ideal_ms"] for r in rows if "ideal_ms" in r)
print(f"conv blocks: measured {tot:.2f} ms, per-layer roofline (max of 6.4 TB/s HBM passes as launched, 1400 TF/s) {ideal:.2f} ms")
by_res = {}
for r in rows:
    if "ideal_ms" in r:
        k = (r["phase"], r["Ho"], r["Wo"])
        a = by_res.setdefault(k, [0.0, 0.0, 0])
        a[0] += r["ms"]; a[1] += r["ideal_ms"]; a[2] += 1
for k in sorted(by_res):
    a = by_res[k]
    print(f"  {k[0]} {k[1]:4d}x{k[2]:<4d} n={a[2]:3d}  measured {a[0]:7.3f} ms  roofline {a[1]:7.3f} ms  x{a[0] / a[1]:.2f}")
for name, ts, fns in (("fwd", ft, plan.fwd), ("bwd", bt, plan.bwd)):
    rows = sorted(zip(ts, [describe(f) for f in fns]), reverse=True)[:18]
    for t, d in rows:
        print(f"  {name} {t:8.3f} ms  {d}")
    agg = {}
    for t, f in zip(ts, fns):
        k = describe(f).split(" ")[0]
        agg[k] = agg.get(k, 0.0) + t
    print("  by kind:", {k: round(x, 2) for k, x in sorted(agg.items(), key=lambda kv: -kv[1])})
