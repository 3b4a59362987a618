#!/usr/bin/env python
"""Per-layer parity of the native plan against the oracle, two views:
  * teacher-forced (default): the oracle recomputes every layer from the native inputs of that layer -> tight, a
    kernel bug shows as a layer with elements over 2 ulp;
  * DYK_PARITY_FP32=1: free-running fp32 oracle -> shows how storage precision drifts with depth.
    python tools/layer_parity.py [cfg] [H] [W] [B] [fp16|bf16]"""
import os, sys
os.environ["DYK_NO_REUSE"] = "1"
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
import bench, models
from oracle import layerwise

cfg = sys.argv[1] if len(sys.argv) > 1 else "kaist_dyolov3_add_sl.cfg"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
W = int(sys.argv[3]) if len(sys.argv) > 3 else 160
B = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dt = torch.bfloat16 if (len(sys.argv) > 5 and sys.argv[5] == "bf16") else torch.float16
path, ref, st = bench.oracle_objects(cfg)
m = models.YOLO(path, (H, W)); m.load_state_dict(st); m = m.cuda().eval()
m.use_cuda_graph = False
m.compute_dtype = dt
g = torch.Generator().manual_seed(3)
v = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=g)
l = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=g)
dual = "second_index" in m.net_info
vf, lf = v.float() / 255, (l.float() / 255 if dual else None)
with torch.no_grad():
    io, p = m(v.cuda(), l.cuda()) if dual else m(v.cuda())
torch.cuda.synchronize()
print(f"{cfg} {H}x{W} B={B} {dt}")
if os.environ.get("DYK_PARITY_FP32", "0") == "1":
    got, _ = layerwise.native_layers(m)
    with torch.no_grad():
        (io_ref, p_ref), every = ref.forward(st, vf, lf, keep_layers=True)
    for i, nat in sorted(got.items()):
        r = every[i]
        err = (nat - r).abs()
        rms = r.pow(2).mean().sqrt().item()
        print(f"L{i:3d} {m.module_defs[i]['type']:14s} C={r.shape[1]:5d} {r.shape[2]:3d}x{r.shape[3]:3d} rms_ref={rms:8.4f} "
              f"rel_rms={err.pow(2).mean().sqrt().item() / (rms + 1e-12):9.2e} max_abs={err.max().item():9.3e}")
    box = (io.cpu()[..., :4] - io_ref[..., :4]).abs() / (io_ref[..., :4].abs() + 8.0)
    print("free-running fp32: box rel max", box.max().item(), "mean", box.mean().item(), "conf max",
          (io.cpu()[..., 4:] - io_ref[..., 4:]).abs().max().item())
else:
    rows = layerwise.compare(m, ref, st, vf, lf, dt)
    for r in rows:
        flag = "  <<< OVER 2 ULP" if r.get("over_2ulp", 1) else ""
        print(f"L{r['layer']:3d} {r['type']:14s} max_diff={r.get('max_diff', -1):9.3e} over_2ulp={r.get('over_2ulp', -1):7d} "
              f"/{r.get('n', 0):9d} frac_diff={r.get('frac_diff', -1):8.2e} rel_rms={r.get('rel_rms', -1):8.2e}{flag}")
    bad = [r for r in rows if r.get("over_2ulp", 1)]
    print(f"teacher-forced: {len(rows)} layers compared, {len(bad)} with elements over 2 ulp, worst frac_diff "
          f"{max(r.get('frac_diff', 0) for r in rows):.2e}")
