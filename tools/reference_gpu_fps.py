#!/usr/bin/env python
"""Measurement infrastructure (like bench.py's cpu_baseline leg): the reference's algorithm — the oracle port, i.e. the same
torch library ops the reference dispatches to, plus the reference NMS loop around torchvision.ops.nms — timed in eager
PyTorch on the SAME B200.  This is the denominator of the north-star's ">= 10x the reference GPU FPS" target; the real
reference cannot be imported on the GPU box.  fp32 (cuDNN TF32 as the reference leaves the default) and autocast fp16.
    python tools/reference_gpu_fps.py [cfg] [batch] [steps]"""
import sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO), str(REPO / "double-yolo-kaist_b200")]
import torch
import bench
from tools.nms_sweep_ref import reference_gpu_nms

cfg = sys.argv[1] if len(sys.argv) > 1 else "kaist_dyolov4_fshare_global_concat_se3.cfg"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
path, ref, st = bench.oracle_objects(cfg)
st = {k: v.cuda() for k, v in st.items()}
dual = "second_index" in ref.net
v8, l8 = [t.cuda() for t in bench.synthetic_frames(B, 0)]


def step(autocast):
    v, l = v8.float() / 255.0, (l8.float() / 255.0 if dual else None)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
        io, _ = ref.forward(st, v, l)
    return reference_gpu_nms(io.float(), 0.01, 0.6)


for autocast in (False, True):
    for _ in range(3):
        step(autocast)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step(autocast)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{cfg} bs{B} reference-on-GPU ({'autocast fp16' if autocast else 'fp32 (TF32 convs)'}): "
          f"{B * steps / dt:8.1f} paired frames/s  ({dt / steps * 1e3:.1f} ms/step, forward + NMS, cudnn.benchmark=False)")

# training step (BASELINE config 3): forward in train mode + backward through autograd + SGD, as the reference's loop does
# (train_utils/kaist_train_eval_utils.py:74-108) minus compute_loss, which bench.py --mode train also replaces by a probe loss
if "--train" in sys.argv:
    params = {k: t.clone().requires_grad_(True) for k, t in st.items()
              if t.dtype.is_floating_point and not k.endswith(("running_mean", "running_var"))}
    state = dict(st)
    state.update(params)
    opt = torch.optim.SGD(list(params.values()), lr=1e-4, momentum=0.9)
    for dt_name, dt in (("bf16 autocast", torch.bfloat16), ("fp16 autocast", torch.float16)):
        def train_step():
            v, l = v8.float() / 255.0, (l8.float() / 255.0 if dual else None)
            with torch.autocast("cuda", dtype=dt):
                p = ref.forward(state, v, l, training=True)
            loss = sum((t.float() ** 2).mean() for t in p)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        for _ in range(3):
            train_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            train_step()
        torch.cuda.synchronize()
        dt_s = time.perf_counter() - t0
        print(f"{cfg} bs{B} reference-on-GPU TRAIN ({dt_name}): {B * steps / dt_s:8.1f} paired frames/s  "
              f"({dt_s / steps * 1e3:.1f} ms/step, forward + backward + SGD)")
