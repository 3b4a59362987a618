#!/usr/bin/env python
"""End-to-end drift of 16-bit storage, native vs the reference's own reduced-precision path, on identical frames/weights.

For each BASELINE model at 512x640: the fp32 CPU oracle is the ground truth; against it are measured (same metrics)
  * native fp16 / bf16 (models.YOLO, tcgen05 kernels, 16-bit NHWC storage, fp32 accumulation),
  * the oracle run on the GPU under torch.autocast(fp16 / bf16)  == what the reference's modules do under autocast
    (train_utils/kaist_train_eval_utils.py:74; cuDNN kernels, 16-bit storage),
  * the oracle on the GPU in "fp32" with PyTorch's default TF32 convolutions == what the reference's evaluate.py runs on
    an Ampere+ GPU,
  * native fp32 mode (3-way bf16 split products).
    python tools/drift_table.py [out.json] [B]"""
import json
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
import models
from dyk import cfg_zoo
from oracle import darknet_ref as dr
from oracle import weights as ow

out = sys.argv[1] if len(sys.argv) > 1 else None
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H, W = 512, 640
DEV = "cuda:0"


def metrics(io, p, io_ref, p_ref):
    io, io_ref = io.float().cpu(), io_ref.float().cpu()
    rms = [float((a.float().cpu() - b.float().cpu()).pow(2).mean().sqrt()) for a, b in zip(p, p_ref)]
    box = (io[..., :4] - io_ref[..., :4]).abs() / (io_ref[..., :4].abs() + 8.0)
    return dict(logit_rms=max(rms), box_mean=float(box.mean()), conf_max=float((io[..., 4:] - io_ref[..., 4:]).abs().max()),
                box_rel_max=float(((io[..., :4] - io_ref[..., :4]).abs() / io_ref[..., :4].abs().clamp_min(1.0)).max()))


table = {}
for name in sorted(cfg_zoo.ZOO):
    path = cfg_zoo.materialize(name)
    ref = dr.DarknetRef(path)
    st = ow.make_calibrated_state(ref, seed=0)
    dual = "second_index" in ref.net
    g = torch.Generator().manual_seed(7)
    v = torch.rand((B, 3, H, W), generator=g)
    l = torch.rand((B, 3, H, W), generator=g) if dual else None
    with torch.no_grad():
        io32, p32 = ref.forward(st, v, l)
    st_gpu = {k: t.to(DEV) for k, t in st.items()}
    vg, lg = v.to(DEV), (l.to(DEV) if dual else None)
    row = {}
    for dt, tag in ((torch.float16, "fp16"), (torch.bfloat16, "bf16")):
        with torch.no_grad(), torch.autocast("cuda", dtype=dt):
            io_a, p_a = ref.forward(st_gpu, vg, lg)
        row[f"reference_autocast_{tag}"] = metrics(io_a, p_a, io32, p32)
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        with torch.no_grad():
            io_a, p_a = ref.forward(st_gpu, vg, lg)
        row["reference_gpu_tf32" if tf32 else "reference_gpu_fp32"] = metrics(io_a, p_a, io32, p32)
    m = models.YOLO(path, (H, W))
    m.load_state_dict(st, strict=True)
    m = m.to(DEV).eval()
    for dt, tag in ((torch.float16, "fp16"), (torch.bfloat16, "bf16"), (torch.float32, "fp32")):
        m.compute_dtype = dt
        with torch.no_grad():
            io_n, p_n = m(vg, lg) if dual else m(vg)
        torch.cuda.synchronize()
        row[f"native_{tag}"] = metrics(io_n, p_n, io32, p32)
    table[name] = row
    print(name)
    for k, r in row.items():
        print(f"  {k:28s} " + "  ".join(f"{a}={b:.3e}" for a, b in r.items()))
    del m
    torch.cuda.empty_cache()
if out:
    Path(out).write_text(json.dumps(table, indent=1))
