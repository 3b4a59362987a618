#!/usr/bin/env python
"""Headline benchmark: paired RGB+LWIR 640x512 frames/s through the dual-stream YOLO hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--cfg NAME] [--batch B]

A step = one batch of synthetic paired frames through ``model(visible, lwir)`` (eval) followed by the batched
``non_max_suppression`` — what evaluate.py does per batch (reference evaluate.py:70-73).  Workload at N=1 is
BASELINE.json configs[1]: kaist_dyolov3_add_sl.cfg, 512x640 (HxW), batch 16, fp16.  With N>1 (torchrun, one
rank per GPU) every rank runs the same per-GPU batch on its own shard of frames — the path has no exchange
step in inference, so scaling is "weak" and there is no data-path collective; only the timing is reduced
(max over ranks).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` goes through the public API
with pinned host uint8 frames (H2D inside the timed region, detections read back), `roofline` is for the
tcgen05 convolution kernel family (FLOPs of all dense convs in the step / their summed CUDA-event time),
`cpu_baseline` is the oracle port timed on this box's host cores on a bounded sample.
`--impl reference` times that CPU port alone (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
PKG = REPO / "double-yolo-kaist_b200"
for p in (str(REPO), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "paired RGB+LWIR 640x512 frames/sec (forward + batched NMS, eval)"
UNIT = "frames/s"
H, W = 512, 640
CONF, IOU = 0.01, 0.6   # evaluate.py:73
FALLBACK_PEAK_TFLOPS, FALLBACK_PEAK_GBS = 1590.0, 6650.0   # B200_PROFILING.md fallback


def peaks():
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            d = json.loads(f.read_text())
            return float(d.get("bf16_tflops", FALLBACK_PEAK_TFLOPS)), float(d.get("hbm_gbs", FALLBACK_PEAK_GBS)), "measured"
        except Exception:  # noqa: BLE001
            pass
    return FALLBACK_PEAK_TFLOPS, FALLBACK_PEAK_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def mark(self):
        """Call when the timed region starts: only samples taken from here on are reported.  (The process itself is
        started before the warm-up: nvidia-smi's start-up attaches to the driver and was seen to stall kernel launches
        for tens of milliseconds when it coincided with the timed region.)"""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = self.rows[self.first:]
        if not rows:                   # timed region shorter than one sampling period: keep the GPU busy state's last sample
            time.sleep(0.06)
            rows = self.rows[-1:]
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_frames(B, seed):
    g = torch.Generator().manual_seed(seed)
    v = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=g)
    l = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=g)
    return v, l


def oracle_objects(cfg_name):
    from dyk import cfg_zoo
    from oracle import darknet_ref as dr
    from oracle import weights as ow
    path = cfg_zoo.materialize(cfg_name)
    ref = dr.DarknetRef(path)
    st = ow.make_calibrated_state(ref, seed=0)
    return path, ref, st


def time_cpu_port(ref, st, frames_per_step, steps, warmup, budget_s=25.0):
    """The reference's algorithm (oracle port: same torch CPU ops the reference dispatches to + numpy NMS
    restatement) on the host cores.  Returns (frames/s, steps actually timed)."""
    from oracle import nms_ref
    dual = "second_index" in ref.net
    v8, l8 = synthetic_frames(frames_per_step, seed=100)
    v, l = v8.float() / 255.0, (l8.float() / 255.0 if dual else None)

    def step():
        with torch.no_grad():
            io, _ = ref.forward(st, v, l)
        nms_ref.non_max_suppression(io.numpy(), CONF, IOU, multi_label=False)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return frames_per_step * done / dt, done, dt


def run_reference_arm(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    _, ref, st = oracle_objects(args.cfg)
    fps, done, dt = time_cpu_port(ref, st, args.ref_frames, args.steps, min(args.warmup, 1), budget_s=120.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": min(args.warmup, 1), "ms_per_step": dt / done * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.cfg} {W}x{H} eval forward + NMS", "batch_per_step": args.ref_frames,
                   "note": "CPU port of the reference path (same torch CPU ops); bounded sample of the bs-16 workload"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{done} steps of {args.ref_frames} paired frames, fp32, torch threads={os.cpu_count()}"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def conv_breakdown(plan, x, y, iters=3):
    """Per-launch CUDA-event timing of the captured plan run eagerly: time and FLOPs of the tcgen05 conv
    launches and total of everything else (each launch is bracketed by events on the launching stream)."""
    from dyk.plan import _ConvStep, _Call
    per = []
    for it in range(iters + 1):
        plan.run_stems(x, y)
        # keep the device busy (~20 ms) while the host enqueues the ~140 bracketed launches: each launch costs ~18 us of host
        # work (ctypes + tensor-map encoding), and with an empty queue that host time would sit inside the event bracket
        # (tools/chain_bench.py: 58.9 us eager bracket vs 40.8 us per launch inside a graph for the same kernel)
        torch.cuda._sleep(40_000_000)
        evs = []
        for s in plan.steps:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            s()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        if it == 0:
            continue
        per.append([a.elapsed_time(b) for a, b in evs])
    mean = [sum(col) / len(col) for col in zip(*per)]
    conv_ms = sum(t for t, s in zip(mean, plan.steps) if isinstance(s, _ConvStep))
    other_ms = sum(t for t, s in zip(mean, plan.steps) if not isinstance(s, _ConvStep))
    flops = 0.0
    n_conv = 0
    dom = {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0}      # launches that dispatch to conv3x3_halo2_kernel
    for t, s in zip(mean, plan.steps):
        if isinstance(s, _ConvStep):
            n_conv += 1
            k = s.kw["k"]
            up = 4 if s.kw["upsample2x"] else 1
            opix = s.y.N * s.y.H * s.y.W // up
            cout = s.e["conv"].out_channels
            fl = 2.0 * opix * cout * s.x.C * k * k
            flops += fl
            if uses_halo2(s):
                dom["ms"] += t
                dom["flops"] += fl
                dom["n"] += 1
                res = 1 if s.kw.get("res") is not None else 0
                srcs = 2 if s.kw.get("x2") is not None else 1          # dual-source modality fusion reads both tensors
                dom["bytes"] += 2.0 * (srcs * s.x.N * s.x.H * s.x.W * s.x.C + opix * cout * (1 + res) + cout * s.x.C * k * k)
    # depthwise launches (dwconv_tile_kernel): HBM-bound, algorithmic bytes = (in + out) * 2 B
    from dyk import ops as _ops
    dw = {"ms": 0.0, "bytes": 0.0, "n": 0}
    for t, s in zip(mean, plan.steps):
        if isinstance(s, _Call) and s.fn is _ops.nhwc_dwconv:
            xin, yout = s.args[0], s.args[4]
            dw["ms"] += t
            dw["bytes"] += 2.0 * (xin.N * xin.H * xin.W * xin.C + yout.N * yout.H * yout.W * yout.C)
            dw["n"] += 1
    dom["dw"] = dw
    return conv_ms, other_ms, flops, n_conv, dom


def uses_halo2(step):
    """Mirror of the dispatch rule in csrc/conv_halo2.cu (conv3x3_halo2_try): 3x3 stride-1 pad-1 layers whose spatial size
    fills >= 60 % of the 8x16-pixel sub-tiles and that have either >= 256 stored output channels and >= 64 input channels
    (streamed weights) or <= 64 input channels and 64 / 96 / 128 output channels (resident weights)."""
    kw = step.kw
    if os.environ.get("DYK_HALO2", "1") == "0" or os.environ.get("DYK_NO_HALO", "0") == "1":
        return False
    if not (kw["k"] == 3 and kw["stride"] == 1 and kw["pad"] == 1 and not kw["upsample2x"] and not kw.get("out_f32", False)):
        return False
    x, y = step.x, step.y
    subs = -(-x.W // 8) * -(-x.H // 16)
    if x.W * x.H / (subs * 128) < 0.6:
        return False
    if y.C >= 256 and x.C >= 64:
        return True
    resident = os.environ.get("DYK_HALO2_RES", "1") != "0" and kw.get("x2") is None
    return resident and 16 <= x.C <= 64 and 64 <= y.C <= 128 and y.C % 32 == 0 and subs * x.N >= 148


def conv_flops_per_frame(model, Hh, Ww):
    """2 * sum_conv Cout*Ho*Wo*Cin/groups*k*k for one frame (pair), from the module list."""
    from dyk import plan as P
    ops_, _, _, _ = P.build_ops(model, Hh, Ww, "second_index" in model.net_info)
    fl = 0.0
    for op in ops_:
        if isinstance(op, P.ConvOp):
            c = op.conv
            fl += 2.0 * op.out.H * op.out.W * c.out_channels * (c.in_channels // c.groups) * c.kernel_size[0] * c.kernel_size[1]
    return fl


def run_train(args, rank, local, world, dev, cfg=None, dtype=None, steps=None, warmup=None, sampler=None):
    """BASELINE configs[2]: one optimizer step = forward (batch-statistics BN) + compute_loss on synthetic labels + backward
    + gradient all-reduce over NCCL + SGD(momentum, nesterov) update, batch 16 per GPU, bf16 storage / fp32 accumulation.
    The loss is the native compute_loss (reference build_utils/utils.py:209-384) with the reference's
    hyp.scratch.4.yaml weights; labels as SURVEY §8d: 3 boxes per frame, cls 0, xy~U(.1,.9), w~U(.03,.1), h~U(.1,.3)."""
    import torch.distributed as dist
    import models
    from dyk import _native as nat
    from build_utils.utils import compute_loss
    from dyk import cfg_zoo, dist_utils

    cfg = cfg or args.cfg
    dtype = dtype or args.dtype
    steps = steps or args.steps
    warmup = warmup or args.warmup
    path = cfg_zoo.materialize(cfg)
    torch.manual_seed(0)
    model = models.YOLO(path, (H, W)).to(dev).train()
    model.compute_dtype = torch.float16 if dtype == "fp16" else torch.bfloat16
    dual = "second_index" in model.net_info
    B = args.batch
    params = [p for p in model.parameters() if p.requires_grad]
    from dyk import optim as dyk_optim
    # train.py:85-91: optim.SGD(pg, lr, momentum, weight_decay, nesterov=True) — here its fused multi-tensor replacement
    # (one native launch for all 568 tensors); DYK_TORCH_OPT=1 times torch.optim.SGD(foreach=True) instead
    if os.environ.get("DYK_TORCH_OPT", "0") == "1":
        opt = torch.optim.SGD(params, lr=1e-4, momentum=0.937, weight_decay=5e-4, nesterov=True, foreach=True)
    else:
        opt = dyk_optim.FusedSGD(params, lr=1e-4, momentum=0.937, weight_decay=5e-4, nesterov=True)
    ring = 2
    host = [tuple(t.pin_memory() for t in synthetic_frames(B, seed=1000 * rank + i)) for i in range(ring)]
    resident = [(v.to(dev), l.to(dev)) for v, l in host]
    model.nc, model.gr = 1, 1.0
    # gradient all-reduce overlapped with the backward pass (dyk.dist_utils.OverlappedAllReduce); DYK_OVERLAP=0 = after it
    overlap = world > 1 and os.environ.get("DYK_OVERLAP", "1") != "0"
    comm = os.environ.get("DYK_GRAD_COMM", "fp32")          # "bf16": gradients cross NVLink as bf16 (half the bytes)
    if overlap:
        model.grad_reducer = dist_utils.OverlappedAllReduce(comm_dtype=torch.bfloat16 if comm == "bf16" else None)
    model.hyp = {"box": 3.54, "cls": 37.4, "obj": 64.3, "cls_pw": 1.0, "obj_pw": 1.0, "iou_t": 0.20, "fl_gamma": 0.0}
    if "yolov4" in model.cfg:
        model.hyp["ciou"] = 1.0
    g = torch.Generator().manual_seed(1 + rank)
    nt = 3 * B
    labels = torch.zeros((nt, 6))
    labels[:, 0] = torch.arange(nt) // 3
    labels[:, 2:4] = torch.rand((nt, 2), generator=g) * 0.8 + 0.1
    labels[:, 4] = torch.rand((nt,), generator=g) * 0.07 + 0.03
    labels[:, 5] = torch.rand((nt,), generator=g) * 0.2 + 0.1
    labels = labels.to(dev)

    def step(v, l):
        p = model(v, l) if dual else model(v)
        parts = compute_loss(p, labels, model)
        loss = parts["box_loss"] + parts["obj_loss"] + parts["class_loss"]
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if not overlap:
            dist_utils.allreduce_gradients(params)
        opt.step()
        return loss

    def step_resident(i):
        return step(*resident[i % ring])

    def step_e2e(i):
        v, l = host[i % ring]
        loss = step(v.to(dev, non_blocking=True), l.to(dev, non_blocking=True))
        return float(loss.item())          # the step's result read back

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        return dist_utils.max_over_ranks(e0.elapsed_time(e1), dev)

    own_sampler = sampler is None
    if own_sampler:
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
    for i in range(warmup):
        step_resident(i)
    step_e2e(0)
    torch.cuda.synchronize()
    if own_sampler:
        sampler.mark()
    l0 = nat.launch_count()
    ms = timed(step_resident, steps)
    launches = nat.launch_count() - l0
    clocks = sampler.stop() if (rank == 0 and own_sampler) else None
    ms_e2e = timed(step_e2e, steps)
    # the collective of the path, timed alone on the same flat gradient buffer (device events, max over ranks)
    ar_ms = None
    plan = model._train_plans.last_plan
    if world > 1:
        buf = plan.flat
        for _ in range(2):
            dist.all_reduce(buf, op=dist.ReduceOp.AVG)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dist.all_reduce(buf, op=dist.ReduceOp.AVG)
        e1.record()
        barrier()
        ar_ms = dist_utils.max_over_ranks(e0.elapsed_time(e1), dev) / 5
    line = None
    if rank == 0:
        peak_tf, peak_gb, peak_kind = peaks()
        sustained = peak_tf
        try:
            sustained = float(json.loads((REPO / "MEASURED_PEAKS.json").read_text()).get("bf16_tflops_sustained", peak_tf))
        except Exception:  # noqa: BLE001
            pass
        fl_step = 3.0 * conv_flops_per_frame(model, H, W) * B          # fwd + dgrad + wgrad
        achieved = fl_step / (ms / steps / 1e3) / 1e12
        red = getattr(model, "grad_reducer", None)
        line = {
            "metric": "paired RGB+LWIR 640x512 frames/sec (training step: forward + backward + all-reduce + SGD)",
            "value": world * B * steps / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": f"{cfg} {W}x{H} train step, batch {B}/GPU, default-initialised weights, native "
                                   "compute_loss (CIoU + objectness BCE) on 3 synthetic labels per frame, fused SGD (momentum 0.937, nesterov, weight decay 5e-4)", "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"dp{world} (replicas; gradient all-reduce in flat buckets"
                                      + (" overlapped with the backward pass)" if overlap else " after the backward pass)"),
                       "l2": "activations saved for backward (tens of GB) exceed the 126 MB L2; no explicit flush"},
            "clocks": clocks,
            "e2e": {"value": world * B * steps / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": 2 * B * 3 * H * W, "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "cuda_graphs": {"forward": plan._fwd_graph is not None,
                            "backward_segments": len(plan._bwd_graphs) if plan._bwd_graphs else 0},
            "allreduce": {"collective": "NCCL all-reduce (AVG) of the flat gradient buffer" if world > 1 else None,
                          "bytes": plan.grad_numel * (2 if comm == "bf16" else 4) if world > 1 else 0,
                          "wire_dtype": comm if world > 1 else None,
                          "buckets_per_step": (red.calls // max(red.steps, 1)) if red is not None and getattr(red, "steps", 0) else 0,
                          "overlapped_with_backward": bool(overlap),
                          "ms_alone": ar_ms,
                          "note": "ms_alone = the same buffer all-reduced in one call with nothing else running (NCCL kernel "
                                  "time, device events, max over ranks); inside the step the buckets overlap the backward graphs"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                         "frac": achieved / sustained, "traffic": None, "peak_source": peak_kind + " (sustained)",
                         "kernel": "whole training step (3 x forward conv FLOPs / step time)",
                         "algorithmic_gflop_per_step": fl_step / 1e9},
            "train_plan_gb": plan.bytes_allocated / 1e9,
        }
    model._train_plans.invalidate()
    del model, opt
    torch.cuda.empty_cache()
    return line


TRAFFIC_PROFILE = "profiles/r02_dram_traffic_dyolov3.txt"


def _profile_traffic(families):
    """Sum of `read + write` GB of the named kernel families in the committed ncu DRAM capture of one bs-16 dyolov3_add_sl
    step (profiles/r02_dram_traffic_dyolov3.txt, written by tools/dram_summary.py from `ncu --metrics dram__bytes_read.sum,
    dram__bytes_write.sum` over tools/one_forward.py) -> (bytes, launches); (None, 0) when the file is missing."""
    import re
    path = REPO / TRAFFIC_PROFILE
    if not path.exists():
        return None, 0
    total, n = 0.0, 0
    for ln in path.read_text().splitlines():
        m = re.match(r"\s*(\S+)\s+n=\s*(\d+)\s+[\d.]+ ms\s+read\s+([\d.]+) GB\s+write\s+([\d.]+) GB", ln)
        if m and m.group(1) in families:
            total += (float(m.group(3)) + float(m.group(4))) * 1e9
            n += int(m.group(2))
    return (total, n) if n else (None, 0)


def conv_dram_traffic(cfg, B):
    """DRAM bytes moved by all convolution launches of one step, from the committed ncu capture of this exact workload; None
    for workloads that were not captured.  It is below the algorithmic 11.04 GB because the 126 MB L2 keeps part of each
    layer's output for its consumer."""
    if cfg == "kaist_dyolov3_add_sl.cfg" and B == 16:
        return _profile_traffic({"conv3x3_halo2_kernel", "conv_tc_kernel", "conv3x3_halo_kernel", "stem_tc_kernel"})[0]
    return None


def halo2_dram_traffic(cfg, B):
    """DRAM bytes of the conv3x3_halo2_kernel launches of one step from the same capture."""
    if cfg == "kaist_dyolov3_add_sl.cfg" and B == 16:
        return _profile_traffic({"conv3x3_halo2_kernel"})[0]
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400, help="timed steps (default: ~2 s of device time at 5.4 ms per step)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer = BASELINE configs[1] (headline, default); train = configs[2]: dyolov4_fshare bf16 "
                         "forward + backward + gradient all-reduce + SGD step")
    ap.add_argument("--cfg", default=None)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--dtype", default=None, choices=["fp16", "bf16"])
    ap.add_argument("--ref-frames", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-leg", action="store_true", help="skip the short configs[2] training leg of the default run")
    ap.add_argument("--train-steps", type=int, default=12)
    ap.add_argument("--sustain-s", type=float, default=2.0,
                    help="after the K timed steps, keep running the same step for about this long (clock samples, "
                         "thermally settled throughput); reported under 'sustained', never as 'value'")
    args = ap.parse_args()

    if args.cfg is None:
        args.cfg = "kaist_dyolov3_add_sl.cfg" if args.mode == "infer" else "kaist_dyolov4_fshare_global_concat_se3.cfg"
    if args.dtype is None:
        args.dtype = "fp16" if args.mode == "infer" else "bf16"
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import models
    from build_utils.utils import nms_raw
    from dyk import _native as nat

    if args.mode == "train":
        line = run_train(args, rank, local, world, dev)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    path, ref, st = oracle_objects(args.cfg)
    model = models.YOLO(path, (H, W))
    model.load_state_dict(st, strict=True)
    model = model.to(dev).eval()
    model.compute_dtype = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    dual = "second_index" in model.net_info
    B = args.batch

    # distinct frames per rank and per step slot (a ring of host batches; pinned for the e2e leg)
    ring = 4
    host = [tuple(t.pin_memory() for t in synthetic_frames(B, seed=1000 * rank + i)) for i in range(ring)]
    resident = [(v.to(dev), l.to(dev)) for v, l in host]

    # The evaluate.py-style loop through the repo's public pipeline helper (dyk/pipeline.py): forward of batch i+1
    # overlaps the batched NMS of batch i (second stream) and, in the e2e leg, the upload of batch i+2 (copy stream).
    # Every batch's forward, NMS (and H2D / D2H) lies inside the timed region; `flush` drains the last one.
    from dyk.pipeline import EvalPipeline
    pipe = EvalPipeline(model, CONF, IOU, multi_label=False)
    pipe_h = EvalPipeline(model, CONF, IOU, multi_label=False, to_host=True)    # e2e: detections land in pinned host memory

    inflight = []

    def step_resident(i):
        # the host stays at most three batches ahead of the device (as a loop that consumes its results would): running
        # 100 batches ahead makes the caching allocator hand out fresh blocks for every batch's outputs (the record_stream'ed
        # ones are not reusable yet) and the loop becomes cudaMalloc-bound on the host (7.75 ms per step, both models alike)
        v, l = resident[i % ring]
        out = pipe.submit(v, l if dual else None)
        ev = torch.cuda.Event()
        ev.record()
        inflight.append(ev)
        if len(inflight) > 3:
            inflight.pop(0).synchronize()
        return out

    def step_e2e(i):
        if i == 0 or pipe_h._staged is None:
            pipe_h.stage(*host[i % ring]) if dual else pipe_h.stage(host[i % ring][0])
        res = pipe_h.submit()                    # host tensors with the detections of the previous batch (the step's result)
        nxt = host[(i + 1) % ring]
        pipe_h.stage(*nxt) if dual else pipe_h.stage(nxt[0])  # upload of the next batch overlaps this batch's compute
        if res is not None:
            return int(res[1].sum())             # the host consumes the result
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    debug = os.environ.get("DYK_BENCH_DEBUG") == "1"

    def timed(fn, steps, d2h=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        marks, host = [], []
        for i in range(steps):
            t0 = time.perf_counter()
            fn(i)
            if debug:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append(ev)
                host.append((time.perf_counter() - t0) * 1e3)
        if debug and rank == 0:
            torch.cuda.synchronize()
            dev_ms = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
            print("DEBUG", fn.__name__, "host ms/step", [round(h, 2) for h in host], "device ms between step ends",
                  [round(d, 2) for d in dev_ms], file=sys.stderr, flush=True)
        last = (pipe_h if d2h else pipe).flush()   # the last batch's NMS (and read-back) belongs to the timed region
        if d2h and last is not None:
            int(last[1].sum())
        pipe._staged = None
        pipe_h._staged = None
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # W is a minimum: the warm-up continues until ~0.5 s of steps have run, so that plan building, graph capture and the
    # clock ramp-up of a fresh box lie before the timed region even when the driver asks for W = 5 (40 ms)
    t_w0 = time.perf_counter()
    i = 0
    while i < args.warmup or (time.perf_counter() - t_w0 < 0.5 and i < 400):
        step_resident(i)
        i += 1
        if i % 8 == 0:
            torch.cuda.synchronize()
    pipe.flush()
    for i in range(max(args.warmup, 8)):
        step_e2e(i)
    pipe_h.flush()
    pipe._staged = None
    pipe_h._staged = None
    torch.cuda.synchronize()
    # back to the device-resident loop for a few untimed steps: the first resident steps after the e2e warm-up were seen (one
    # box in four) to run host-bound at ~7.8 ms per step for both models alike while the allocator pools of the three
    # streams re-balance; the same loop is fast again in the sustained leg
    for i in range(8):
        step_resident(i)
    pipe.flush()
    torch.cuda.synchronize()
    sampler.mark()
    l0 = nat.launch_count()
    ms = timed(step_resident, args.steps)
    launches = nat.launch_count() - l0
    ms_e2e = timed(step_e2e, args.steps, d2h=True)
    # sustained leg: the same device-resident step for ~sustain_s seconds, so that the clock / throttle samples cover a
    # region long enough to mean something when the driver asks for a short K (20 steps = 0.11 s)
    sus_steps = max(args.steps, int(args.sustain_s * 1e3 / max(ms / args.steps, 1e-3)))
    ms_sus = timed(step_resident, sus_steps) if args.sustain_s > 0 else None
    clocks = sampler.stop() if rank == 0 else None

    value = world * B * args.steps / (ms / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        plan = model._plans.last_plan
        v, l = resident[0]
        conv_ms, other_ms, flops, n_conv, dom = conv_breakdown(plan, v, l if dual else None)
        peak_tf, peak_gb, peak_kind = peaks()
        achieved_all = flops / (conv_ms / 1e3) / 1e12
        dwk, no_pair_kernel = dom.get("dw") or {"n": 0, "ms": 0.0}, not dom["n"]
        if dom["n"]:
            achieved = dom["flops"] / (dom["ms"] / 1e3) / 1e12
        else:                      # models without a CTA-pair layer (MobileNet): all dense convs together
            achieved, dom = achieved_all, {"ms": conv_ms, "flops": flops, "bytes": None, "n": n_conv}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"{args.cfg} {W}x{H} eval forward + batched NMS, batch {B}/GPU, seeded calibrated "
                                   f"random weights", "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"dp{world} (independent shards, no data-path collective)",
                       "l2": "per-step working set (231 MB weights + >5 GB activations) exceeds the 126 MB L2; "
                             "4 distinct input batches cycled; no explicit flush",
                       "pipeline": "dyk.pipeline.EvalPipeline: NMS of batch i on a second stream overlaps the forward of "
                                   "batch i+1; e2e also uploads batch i+1 on a copy stream and reads the detections back to pinned host memory "
                                   "on the NMS stream; all inside the timed region",
                       "conf_thres": CONF, "iou_thres": IOU},
            "clocks": clocks,
            "sustained": None if ms_sus is None else {
                "value": world * B * sus_steps / (ms_sus / 1e3), "unit": UNIT, "steps": sus_steps, "seconds": ms_sus / 1e3,
                "note": "same device-resident step repeated after the K timed steps; `clocks` covers the timed steps, the "
                        "e2e leg and this leg"},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": 2 * B * 3 * H * W, "d2h_bytes_per_step": B * 100 * 6 * 4 + B * 4},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf, "traffic": halo2_dram_traffic(args.cfg, B) if dom["bytes"] else None,
                         "traffic_source": {"from_profile": TRAFFIC_PROFILE, "measured_in_this_run": False},
                         "peak_source": peak_kind,
                         "kernel": "conv3x3_halo2_kernel (tcgen05 cta_group::2 implicit GEMM: the 3x3 stride-1 layers with >= 256 "
                                   "output channels, the dual-source modality-fusion convs and the resident-weights early "
                                   "layers): the dominant kernel of the step by time",
                         "note": "a launch = one layer, so figures are sums over the kernel's launches of one step: "
                                 "achieved = algorithmic FLOPs / summed CUDA-event durations (each launch bracketed by events on "
                                 "the launching stream in an eager single-stream pass that is enqueued behind a ~20 ms device-side "
                                 "sleep, so the brackets hold device time, not the host's per-launch work); traffic = "
                                 "ncu dram__bytes_read.sum + dram__bytes_write.sum of the same launches "
                                 f"({TRAFFIC_PROFILE})",
                         "launches_per_step": dom["n"], "kernel_ms_per_step": dom["ms"],
                         "algorithmic_gflop_per_step": dom["flops"] / 1e9, "algorithmic_bytes_per_step": dom["bytes"],
                         "all_dense_convs": {"achieved": achieved_all, "frac": achieved_all / peak_tf, "launches_per_step": n_conv,
                                             "ms_per_step": conv_ms, "algorithmic_gflop_per_step": flops / 1e9,
                                             "traffic": conv_dram_traffic(args.cfg, B)},
                         "other_kernels_ms_per_step": other_ms,
                         "timing_note": "kernel_ms_per_step / all_dense_convs.ms_per_step are sums of per-launch event "
                                        "brackets of an EAGER single-stream pass; the timed step replays a two-lane CUDA "
                                        "graph in which the two backbones overlap, so those sums can exceed ms_per_step"},
            "cuda_graph": plan.graph is not None,
        }
        if dwk["n"] and (no_pair_kernel or dwk["ms"] > dom["ms"]):
            # MobileNet backbones: the depthwise kernel is the dominant kernel by time and it is HBM-bound
            gbs = dwk["bytes"] / (dwk["ms"] / 1e3) / 1e9
            line["roofline"] = {
                "bound": "hbm", "achieved": gbs, "peak": peak_gb, "unit": "GB/s", "frac": gbs / peak_gb, "traffic": None,
                "peak_source": peak_kind,
                "kernel": "dwconv_tile_kernel (TMA-staged depthwise 3x3 / 5x5, csrc/dwconv_tile.cu): the dominant kernel of the "
                          "step by time on the MobileNet backbones",
                "note": "sums over the kernel's launches of one step: achieved = algorithmic bytes ((in + out) * 2 B per "
                        "launch) / summed CUDA-event durations of an eager single-stream pass enqueued behind a device-side "
                        "sleep; ncu of the same kernel (profiles/r02_ncu_mnv3_kernels.txt): DRAM bytes = algorithmic",
                "launches_per_step": dwk["n"], "kernel_ms_per_step": dwk["ms"], "algorithmic_bytes_per_step": dwk["bytes"],
                "all_dense_convs": {"achieved": achieved_all, "unit": "TFLOP/s", "frac_of_tensor_peak": achieved_all / peak_tf,
                                    "launches_per_step": n_conv, "ms_per_step": conv_ms,
                                    "algorithmic_gflop_per_step": flops / 1e9},
                "other_kernels_ms_per_step": other_ms - dwk["ms"]}
        if not args.no_cpu_baseline and world == 1:      # contract: rank 0 at N = 1 only (at N > 1 the other ranks would spin)
            torch.set_num_threads(os.cpu_count())
            fps, done, dt = time_cpu_port(ref, st, args.ref_frames, 40, 1, budget_s=15.0)   # ~15 s of host work
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{done} steps of {args.ref_frames} paired frames (fp32 oracle port, "
                                              f"torch threads={os.cpu_count()}, {dt:.1f} s)"}
    # configs[2] rides along: a short training leg (dyolov4_fshare, bf16, batch 16 per GPU) so that the driver's N = 1..8
    # records carry the one path that has a collective (gradient all-reduce over NCCL / NVLink).  All ranks take part.
    train = None
    if not args.no_train_leg:
        del pipe, pipe_h
        model._plans.invalidate()
        torch.cuda.empty_cache()
        t = run_train(args, rank, local, world, dev, cfg="kaist_dyolov4_fshare_global_concat_se3.cfg", dtype="bf16",
                      steps=args.train_steps, warmup=3, sampler=ClockSampler(local))
        if rank == 0:
            keep = ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "dtype", "e2e", "gpu_launches", "cuda_graphs",
                    "allreduce", "roofline", "train_plan_gb")
            train = {k: t[k] for k in keep}
            train["config"] = t["config"]
    if rank == 0:
        line["train"] = train
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
