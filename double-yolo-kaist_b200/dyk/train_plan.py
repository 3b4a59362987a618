"""Training-mode execution plan of ``models.YOLO``: forward with batch-statistics BatchNorm and the full backward,
as one ``torch.autograd.Function`` whose body is native launches only.

Replaces what autograd does for the reference (reference train_utils/kaist_train_eval_utils.py:
74-108: ``pred = model(v, l)``, ``scaler.scale(loss).backward()``).  The returned head tensors carry a grad_fn, so the
reference's ``compute_loss`` / ``GradScaler`` / optimizers / DistributedDataParallel work unchanged on top.

Forward (per cfg block, no cross-layer fusion in this mode — every stored tensor is needed by backward):
  [convolutional]+BN : conv -> z (16-bit, pre-BN) | batch statistics (two-stage, deterministic) -> scale/shift,
                       running-stat update | y = act(z*scale + shift)
  heads (no BN)      : conv + bias -> fp32 logits -> permute (p)
  shortcut / route / maxpool / upsample / se : the forward kernels of the eval plan (concats are written in place).
Backward walks the ops in reverse with one gradient buffer per tensor ("first writer overwrites, later writers
accumulate" is resolved when the plan is built):
  BN+act backward (reduce, finalize, apply) -> dz ; wgrad (tcgen05, split-K) ; dgrad (forward conv kernels on rotated
  weights, accumulating through the residual operand) ; axpby routing for shortcut / route ; pool / upsample / SE backward.
Parameter gradients are produced in one flat fp32 buffer (one view per parameter), zeroed once per backward.

Depthwise convolutions (MobileNet backbones, DepthwiseSeparableConv2d) use the CUDA-core kernels of dwconv_bwd.cu; Inception
blocks are their constituent convolutions, a 3x3 max-pool and a concat.
Limits (raise NativeError, never fall back): grouped convolutions that are not depthwise, BatchNorm2d(momentum=None),
shortcuts with more than two operands or unequal widths; a second forward before the backward of the first overwrites the
saved activations (the reference never does that).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _native as nat
from . import ops
from . import plan as P
from . import train_ops as T
from .ops import View

HEAD_PAD = 32   # head logits gradients are padded to a multiple of 32 channels (K of the dgrad GEMM, 16-byte TMA rows)


def head_pad(channels: int) -> int:
    """Padded channel count of a detection head's gradient / data-gradient weights: na*(5+nc) rounded up to 32
    (18 -> 32 for the KAIST cfgs, 75 -> 96 for VOC, 255 -> 256 for COCO)."""
    return (channels + HEAD_PAD - 1) // HEAD_PAD * HEAD_PAD


class _Grad:
    __slots__ = ("view", "written", "alias_of")

    def __init__(self, view):
        self.view, self.written, self.alias_of = view, False, None


def assign_train_lanes(ops_):
    """Two-lane schedule of a training plan (forward and, mirrored, backward).  The dual-stream models are two independent
    backbones between fusion points; their mid / late layers (64x80 and smaller) are far too small to fill 148 SMs, and
    a layer is a chain of 4 (forward) to 6 (backward) dependent launches, so running the LWIR backbone on a second stream
    lets the two chains interleave.  Same greedy rule as dyk.plan.assign_lanes, with the stems as ordinary ops: an op
    follows the lane of the most recent op it depends on, joins on lane 0, and the first op without any ancestor after
    lane 0 has started (the LWIR stem) opens lane 1.  Returns (lane per op, producer op index per Value id)."""
    producer, anc, lanes, last = {}, [], [], [None, None]
    for k, op in enumerate(ops_):
        a = 0
        for v in op.inputs():
            p_ = producer.get(id(v))
            if p_ is not None:
                a |= (1 << p_) | anc[p_]
        anc.append(a)
        out = getattr(op, "out", None)
        if out is not None:
            producer[id(out)] = k
        d0 = last[0] is not None and (a >> last[0]) & 1
        d1 = last[1] is not None and (a >> last[1]) & 1
        if d0 and not d1:
            lane = 0
        elif d1 and not d0:
            lane = 1
        elif d0 and d1:
            lane = 0
        elif a:
            lane = lanes[a.bit_length() - 1]
        else:
            lane = 0 if last[0] is None else 1
        lanes.append(lane)
        last[lane] = k
    return lanes, producer


class _WgradLane:
    """Weight gradients off the critical path.  In the backward pass only  BN-backward(L) -> dgrad(L) -> BN-backward(L-1) ...
    is a dependency chain; the weight gradient of layer L (wgrad kernel + its split-K reduction) only needs dz(L).  It is
    therefore launched on a side stream right after dz(L) exists and runs concurrently with dgrad(L) and the next layers'
    kernels, filling the SMs that the small late layers leave idle.  dz lives in a ring of `nbuf` buffers: before a
    buffer is overwritten the main stream waits for the weight-gradient kernel that last read it.  Inside CUDA-graph
    capture the event record / wait pairs become graph edges (parallel branches).  DYK_WG_SIDE=0 runs everything in order."""

    def __init__(self, device, numel, dtype, nbuf=3):
        self.on = device.type == "cuda" and os.environ.get("DYK_WG_SIDE", "1") != "0"
        self.nbuf = nbuf if self.on else 1
        self.bufs = [torch.empty(numel, dtype=dtype, device=device) for _ in range(self.nbuf)]
        self.pending = [None] * self.nbuf
        self.loose = []
        self.side = torch.cuda.Stream(device=device) if self.on else None
        self.k, self.slot = 0, None

    def next_dz(self, z: View) -> View:
        """dz buffer for the layer whose pre-BN tensor is z (called on the main stream)."""
        i = self.k % self.nbuf
        self.k += 1
        if self.pending[i] is not None:
            torch.cuda.current_stream().wait_event(self.pending[i])
            self.pending[i] = None
        self.slot = i
        return View(self.bufs[i][:z.buf.numel()].view(z.buf.shape), 0, z.C)

    def launch(self, fn, ring=True):
        """Run fn() (weight-gradient launches) after everything enqueued so far on the main stream, on the side stream."""
        if not self.on:
            fn()
            return
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            fn()
            done = torch.cuda.Event()
            done.record(self.side)
        if ring and self.slot is not None:
            self.pending[self.slot] = done
            self.slot = None
        else:
            self.loose.append(done)

    def join(self):
        """Main stream waits for every outstanding weight gradient (end of a graph segment / before gradients are used)."""
        if not self.on:
            return
        main = torch.cuda.current_stream()
        for i, ev in enumerate(self.pending):
            if ev is not None:
                main.wait_event(ev)
                self.pending[i] = None
        for ev in self.loose:
            main.wait_event(ev)
        self.loose = []


class TrainPlan:
    def __init__(self, model, B, H, W, dtype, dual, device, in_dtype=torch.float32, src_hw=None):
        self.model, self.B, self.dtype, self.device, self.dual = model, B, dtype, device, dual
        # Static frame buffers: the caller's batches are copied in (31 MB per uint8 bs-16 pair), so every launch of a step
        # reads fixed addresses and the whole step can be replayed as CUDA graphs.  src_hw != (H, W): multi-scale step, the
        # buffers hold the ORIGINAL frames and the stem kernels (forward and weight gradient) sample their bilinear resize.
        Hs, Ws = src_hw if src_hw is not None else (H, W)
        self.H, self.W = H, W
        self.resize = (Hs, Ws) != (H, W)
        self.in_x = torch.empty((B, 3, Hs, Ws), dtype=in_dtype, device=device)
        self.in_y = torch.empty((B, 3, Hs, Ws), dtype=in_dtype, device=device) if dual else None
        self.ops, self.layer_vals, self.img0, self.img1 = P.build_ops(model, H, W, dual)
        for op in self.ops:
            if isinstance(op, P.ConvOp):
                op.flavor          # grouped (non-depthwise) convolutions raise NativeError here
        P.mark_heads(self.ops)
        P.place_concats(self.ops)
        self.params = list(model.parameters())
        self.offsets, off = {}, 0
        for prm in self.params:
            self.offsets[id(prm)] = off
            off += (prm.numel() + 31) // 32 * 32          # padded so vector kernels may write whole 32-float groups
        self.grad_numel = off
        self.reducer = None      # set by TrainPlanCache from model.grad_reducer (dyk.dist_utils.OverlappedAllReduce)
        self.lanes, self.producer = assign_train_lanes(self.ops)
        self.two_lanes = (device.type == "cuda" and os.environ.get("DYK_TRAIN_LANES", "2") != "1" and 1 in self.lanes)
        if not self.two_lanes:
            self.lanes = [0] * len(self.ops)
        self.lane1 = torch.cuda.Stream(device=device) if self.two_lanes else None
        self._alloc_forward()
        self._bind_forward()
        self._finish_forward_schedule()
        self._bind_backward()
        # gradients of a step are produced in this buffer (one view per parameter) and handed out as a copy
        self._scratch = {}       # kernel workspaces of this plan (addresses are baked into its CUDA graphs)
        self.flat = torch.zeros(self.grad_numel, dtype=torch.float32, device=device)
        self.dps_in = [torch.zeros_like(po) for po in self.p_outs]
        self._sig = None
        self._fwd_graph, self._fwd_warm, self._fwd_launches = None, False, 0
        self._bwd_graphs, self._bwd_key, self._bwd_warm, self._bwd_launches, self._seg_ends = None, None, False, [], []
        self.graph_failed = False

    # ------------------------------------------------------------------------------------------ storage
    def _new(self, Cc, H, W, f32=False):
        dt = torch.float32 if f32 else self.dtype
        return View(torch.empty((self.B, H, W, Cc), dtype=dt, device=self.device), 0, Cc)

    def _alloc_forward(self):
        self.bytes_allocated = 0
        for op in self.ops:
            v = getattr(op, "out", None)
            if v is None or v.place:
                continue
            v.view = self._new(v.C, v.H, v.W, f32=v.f32)
            self.bytes_allocated += v.view.buf.numel() * v.view.buf.element_size()
        for op in self.ops:
            v = getattr(op, "out", None)
            if v is not None and v.place:
                parent, off = v.place
                v.view = View(parent.view.buf, off, v.C)

    def _vec(self, n):
        return torch.empty(n, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------------------------------ forward
    def _bind_forward(self):
        self.fwd = []            # callables taking (x, y)
        self._packed_ops = set()
        self.convs = []          # per ConvOp state (packed weights etc.), refreshed every forward
        self.bns = []
        self.p_outs = []
        self.yolo = []
        dev = self.device
        self.fwd_groups = []     # (op index, first fwd step, one past the last) — the unit of the two-lane schedule
        for op_index, op in enumerate(self.ops):
            if self.fwd_groups:
                self.fwd_groups[-1][2] = len(self.fwd)
            self.fwd_groups.append([op_index, len(self.fwd), len(self.fwd)])
            if isinstance(op, P.ConvOp):
                conv, bn = op.conv, op.bn
                k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
                st = dict(op=op, conv=conv, bn=bn, k=k, s=s, p=p, stem=op.flavor == "stem", dw=op.flavor == "dw")
                if st["dw"] and bn is None:
                    raise nat.NativeError(f"layer {op.layer}: a depthwise convolution without BatchNorm has no training kernels")
                if bn is not None:
                    if bn.momentum is None:
                        raise nat.NativeError("BatchNorm2d(momentum=None) is not supported by the training kernels")
                    if not bn.track_running_stats or not bn.affine:
                        raise nat.NativeError("BatchNorm2d without affine / running stats is not supported")
                    Cc = conv.out_channels
                    st.update(z=self._new(Cc, op.out.H, op.out.W), scale=self._vec(Cc), shift=self._vec(Cc),
                              mean=self._vec(Cc), invstd=self._vec(Cc))
                    self.bytes_allocated += st["z"].buf.numel() * 2
                    self.bns.append(bn)
                elif not op.out.f32:
                    raise nat.NativeError(f"layer {op.layer}: a convolution without BatchNorm must be a detection head")
                if not st["stem"] and not st["dw"] and k * k <= 9 and conv.weight.dtype == torch.float32:
                    # persistent packed copies, refreshed for all layers by one launch per step (_pack_all)
                    O, I = conv.out_channels, conv.in_channels
                    opad = O if bn is not None else head_pad(O)
                    st["w"] = torch.empty((O, k, k, I), dtype=self.dtype, device=dev)
                    st["wd"] = torch.zeros((I, k, k, opad), dtype=self.dtype, device=dev)
                    st["multi"] = True
                    self.bytes_allocated += (st["w"].numel() + st["wd"].numel()) * 2
                self.convs.append(st)
                if st.get("multi"):
                    self._packed_ops.add(op_index)              # reads the packed 16-bit weights (_pack_all)
                self.fwd.append(self._conv_fwd(st))
            elif isinstance(op, P.AddOp):
                self.fwd.extend(self._add_fwd(op))
            elif isinstance(op, P.ConcatOp):
                for s_, off in op.copies:
                    dst = View(op.out.view.buf, op.out.view.c_off + off, s_.C)
                    self.fwd.append(lambda x, y, a=s_.view, b=dst: ops.nhwc_copy(a, b))
            elif isinstance(op, P.PoolOp):
                self.fwd.append(lambda x, y, o=op: ops.nhwc_maxpool(o.src.view, o.out.view, o.k, o.stride))
            elif isinstance(op, P.UpOp):
                self.fwd.append(lambda x, y, o=op: ops.nhwc_upsample(o.src.view, o.out.view, o.s))
            elif isinstance(op, P.SEOp):
                op.pooled = torch.empty((self.B, 32, op.src.C), dtype=torch.float32, device=dev)
                op.gate = torch.empty((self.B, op.src.C), dtype=torch.float32, device=dev)
                self.fwd.append(lambda x, y, o=op: ops.nhwc_se(o.src.view, o.out.view, *ops.se_weights(o.module.fc1, o.module.fc2),
                                                                o.pooled, o.gate))
            elif isinstance(op, P.YoloOp):
                m = op.module
                ny, nx = op.src.H, op.src.W
                if (m.nx, m.ny) != (nx, ny) or m.anchor_vec.device != dev:
                    m.create_grids((nx, ny), dev)
                if not op.src.f32:
                    raise nat.NativeError(f"layer {op.layer}: [yolo] must follow a linear conv without batch norm")
                p_out = torch.empty((self.B, m.na, ny, nx, m.no), dtype=torch.float32, device=dev)
                self.p_outs.append(p_out)
                anchor = m.anchor_vec.to(dev).float().contiguous()
                self.fwd.append(lambda x, y, o=op, po=p_out, an=anchor, mm=m: ops.yolo_decode(
                    o.src.view.buf, o.src.view.stride, po, None, N=self.B, ny=o.src.H, nx=o.src.W, na=mm.na, no=mm.no,
                    anchor_vec=an, stride=mm.stride, v4=mm.bf_type == "yolov4", rows_total=0, row_off=0, in_kind=2))

    def _finish_forward_schedule(self):
        """Cross-lane dependencies of the forward groups: group k waits for the event recorded after group p for every
        input produced on the other lane."""
        if self.fwd_groups:
            self.fwd_groups[-1][2] = len(self.fwd)
        self.fwd_waits, self.fwd_records = {}, set()
        for k, op in enumerate(self.ops):
            for v in op.inputs():
                p_ = self.producer.get(id(v))
                if p_ is not None and self.lanes[p_] != self.lanes[k]:
                    self.fwd_waits.setdefault(k, set()).add(p_)
                    self.fwd_records.add(p_)

    def _conv_fwd(self, st):
        op, conv, bn = st["op"], st["conv"], st["bn"]
        k, s, p = st["k"], st["s"], st["p"]

        def run(x, y):
            if st["stem"]:
                src = self.in_x if op.src is self.img0 else self.in_y
                st["x_in"] = src
                w = conv.weight.detach().float().permute(0, 2, 3, 1).contiguous()
                ops.nhwc_stem(src, w, None, None, st["z"], k=k, stride=s, pad=p, act="linear",
                              resize_to=(self.H, self.W) if self.resize else None)
            elif st["dw"]:
                st["w"] = T.dw_weight(conv)
                ops.nhwc_dwconv(op.src.view, st["w"], None, None, st["z"], k=k, stride=s, pad=p, act="linear")
            else:
                if not st.get("multi"):
                    st["w"] = ops.pack_conv_weight(conv.weight, self.dtype)
                if bn is None:
                    st["bias"] = ops.pad_vec(conv.bias) if conv.bias is not None else None
                    ops.nhwc_conv(op.src.view, st["w"], None, st["bias"], op.out.view, k=k, stride=s, pad=p, act="linear",
                                  out_f32=True, cout=conv.out_channels)
                    return
                ops.nhwc_conv(op.src.view, st["w"], None, None, st["z"], k=k, stride=s, pad=p, act="linear",
                              cout=conv.out_channels)
            T.bn_train_stats(st["z"], bn.weight.detach(), bn.bias.detach(), bn.eps, bn.momentum, bn.running_mean,
                             bn.running_var, st["scale"], st["shift"], st["mean"], st["invstd"])
            T.bn_act_apply(st["z"], st["scale"], st["shift"], op.act, op.out.view)
        return run

    def _add_fwd(self, op):
        m = op.module
        if len(op.others) != 1 or op.others[0].C != op.x.C:
            raise nat.NativeError(f"layer {op.layer}: only two-operand, equal-width shortcuts have training kernels")
        steps = []
        op.wall = None
        if m.weight:
            op.wall = torch.empty(2, dtype=torch.float32, device=self.device)
            steps.append(lambda x, y, o=op, mm=m: ops.fusion_weights(mm.w.detach(), o.wall))
        steps.append(lambda x, y, o=op: ops.nhwc_add(o.x.view, o.others[0].view, o.out.view, o.wall))
        return steps

    def _signature(self):
        """Everything a captured step bakes in that the caller may change between steps: parameter / buffer storage
        addresses (graphs and the packing table hold raw pointers) and the BatchNorm scalars."""
        ptrs = tuple(t.data_ptr() for t in self._sig_tensors)
        return ptrs, tuple((bn.eps, bn.momentum) for bn in self.bns)

    def _prepare(self):
        """Host-side part of a step (never captured): rebuild the weight-packing table and drop the graphs when a
        parameter's storage moved (.to(), load of a new tensor, ...) or a BatchNorm's eps / momentum changed."""
        if not hasattr(self, "_sig_tensors"):
            self._sig_tensors = list(self.model.parameters()) + list(self.model.buffers())
        sig = self._signature()
        if sig == self._sig:
            return
        self._sig = sig
        self._fwd_graph, self._bwd_graphs = None, None
        sts = [st for st in self.convs if st.get("multi")]
        self._pack_n = len(sts)
        if not sts:
            return
        rows, tiles = [], 0
        for st in sts:
            wt = st["conv"].weight
            if not wt.is_contiguous() or wt.dtype != torch.float32:
                raise nat.NativeError("convolution weights must be contiguous float32 parameters")
            O, I, k = wt.shape[0], wt.shape[1], st["k"]
            if st["wd"].shape[3] < O or st["wd"].shape[0] != I or st["w"].shape[0] != O:
                raise nat.NativeError("internal: packed weight buffers do not match the convolution's shape")
            rows.append([wt.data_ptr(), st["w"].data_ptr(), st["wd"].data_ptr(), O, I, k * k, st["wd"].shape[3], tiles])
            tiles += ((O + 31) // 32) * ((I + 31) // 32)
        self._pack_desc = torch.tensor(rows, dtype=torch.int64).to(self.device)
        self._pack_tiles = tiles

    def _pack_all(self):
        """One launch re-packs every dense convolution weight into its forward and data-gradient layouts (the optimizer
        changed them since the last step); the descriptor table is maintained by _prepare."""
        if not self._pack_n:
            return
        nat.call("dyk_pack_weights_multi", ops._p(self._pack_desc), self._pack_n, self._pack_tiles, ops._DT[self.dtype], ops._stream())
        nat.count_launches()

    def _forward_body(self):
        main = torch.cuda.current_stream()
        packed = None
        wg0 = self.wgs[0]
        if wg0.on:
            # the per-step re-packing of all weights (one launch, ~0.9 GB of traffic) runs on a side stream while the
            # stem convolutions — which read the fp32 parameters directly — and their BatchNorm run on the lanes
            fork = torch.cuda.Event()
            fork.record(main)
            wg0.side.wait_event(fork)
            with torch.cuda.stream(wg0.side):
                self._pack_all()
                packed = torch.cuda.Event()
                packed.record(wg0.side)
        else:
            self._pack_all()
        if self.two_lanes:
            fork1 = torch.cuda.Event()
            fork1.record(main)
            self.lane1.wait_event(fork1)
        streams = (main, self.lane1)
        pack_seen = [packed is None, packed is None]
        events = {}
        for k, lo, hi in self.fwd_groups:
            lane = self.lanes[k]
            st = streams[lane]
            if not pack_seen[lane] and k in self._packed_ops:
                st.wait_event(packed)
                pack_seen[lane] = True
            for p_ in self.fwd_waits.get(k, ()):
                st.wait_event(events[p_])
            if lane == 0:
                for i in range(lo, hi):
                    self.fwd[i](self.in_x, self.in_y)
            else:
                with torch.cuda.stream(st):
                    for i in range(lo, hi):
                        self.fwd[i](self.in_x, self.in_y)
            if k in self.fwd_records:
                ev = torch.cuda.Event()
                ev.record(st)
                events[k] = ev
        if self.two_lanes:
            join = torch.cuda.Event()
            join.record(self.lane1)
            main.wait_event(join)
        if not pack_seen[0] and packed is not None:
            main.wait_event(packed)
        if self.bns:
            torch._foreach_add_([bn.num_batches_tracked for bn in self.bns], 1)   # counters, not arithmetic of the path

    def _graphs_on(self):
        return (self.model.use_cuda_graph and not self.graph_failed and self.device.type == "cuda"
                and os.environ.get("DYK_TRAIN_GRAPH", "1") != "0" and not torch.cuda.is_current_stream_capturing())

    def _capture(self, body):
        """Capture `body` (native launches only) into a CUDA graph; on failure fall back to eager launches for good."""
        g = torch.cuda.CUDAGraph()
        try:
            # thread_local: the NCCL watchdog thread may query events while this thread captures
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                body()
            return g
        except Exception as exc:  # noqa: BLE001 - same native kernels, launched one by one
            import warnings
            warnings.warn(f"CUDA graph capture of the training plan failed ({type(exc).__name__}: {exc}); "
                          "running the same native kernels as individual launches")
            self.graph_failed = True
            torch.cuda.synchronize()
            return None

    def forward(self, x, y):
        """One training forward: frames are copied into the plan's static buffers, then the ~900 launches (weight packing,
        convolutions, batch statistics, BN + activation, pools, SE, head permutes) run as ONE graph replay — the first
        call runs them eagerly (one-time kernel attribute setup), the second captures."""
        with T.use_scratch(self._scratch):
            return self._forward(x, y)

    def _forward(self, x, y):
        self.in_x.copy_(x)
        if self.in_y is not None:
            self.in_y.copy_(y)
        self._prepare()
        if self._graphs_on() and self._fwd_warm:
            if self._fwd_graph is None:
                self._fwd_graph = self._capture(self._forward_body)
            if self._fwd_graph is not None:
                self._fwd_graph.replay()
                nat.count_launches(self._fwd_launches)
                return tuple(t.clone() for t in self.p_outs)
        n0 = nat.launch_count()
        self._forward_body()
        self._fwd_launches, self._fwd_warm = nat.launch_count() - n0, True
        return tuple(t.clone() for t in self.p_outs)

    # ------------------------------------------------------------------------------------------ backward
    def _pgrad(self, flat, prm):
        o = self.offsets[id(prm)]
        return flat[o:o + prm.numel()].view(prm.shape)

    def _bind_backward(self):
        """Static schedule of the backward launches.  Each entry is a callable(flat_grads, dps)."""
        consumers = {}
        for op in self.ops:
            for v in op.inputs():
                consumers.setdefault(id(v), []).append(op)
        grads = {}
        self.aux_bwd = []

        def gslot(v):
            g = grads.get(id(v))
            if g is None:
                alias = None
                if v.place is not None:
                    parent, off = v.place
                    cons = consumers.get(id(v), [])
                    if len(cons) == 1 and isinstance(cons[0], P.ConcatOp) and cons[0].out is parent \
                            and all(s is not v for s, _ in cons[0].copies):
                        alias = (parent, off)
                if alias is not None:
                    pg = gslot(alias[0])
                    g = _Grad(View(pg.view.buf, pg.view.c_off + alias[1], v.C))
                    g.alias_of = pg
                elif v.f32:      # head logits: gradient kept as padded 16-bit NHWC
                    g = _Grad(self._new(head_pad(v.C), v.H, v.W))
                else:
                    g = _Grad(self._new(v.C, v.H, v.W))
                    self.bytes_allocated += g.view.buf.numel() * 2
                grads[id(v)] = g
            return g

        def is_written(g):
            return g.written or (g.alias_of is not None and is_written(g.alias_of))

        def root(g):
            while g.alias_of is not None:
                g = g.alias_of
            return g

        cur = {"writes": set(), "reads": set()}      # gradient buffers the group being bound touches (two-lane schedule)

        def claim(v):
            """Returns (grad view of v, accumulate?) and marks it written."""
            g = gslot(v)
            acc = is_written(g)
            g.written = True
            cur["writes"].add(id(root(g)))
            return g.view, acc

        self.grads, self.consumers = grads, consumers
        self.bwd = []
        self.bwd_writes = {}     # index into self.bwd -> parameters whose gradient that step writes
        max_z = max((st["z"].buf.numel() for st in self.convs if "z" in st), default=0)
        self.wgs = [_WgradLane(self.device, max_z, self.dtype) for _ in range(2 if self.two_lanes else 1)]
        self.wg = self.wgs[0]           # the lane the step being executed uses (set per group by _backward_range)
        conv_state = {id(st["op"]): st for st in self.convs}
        yolo_i = len(self.p_outs)
        self.bwd_groups = []            # dict(op, lo, hi, reads, writes): the backward steps of one op, in execution order

        def close_group():
            if cur.get("op") is not None and len(self.bwd) > cur["lo"]:
                self.bwd_groups.append(dict(op=cur["op"], lo=cur["lo"], hi=len(self.bwd), reads=cur["reads"], writes=cur["writes"]))

        for op_index in range(len(self.ops) - 1, -1, -1):
            op = self.ops[op_index]
            close_group()
            cur.update(op=op_index, lo=len(self.bwd), reads=set(), writes=set())
            if isinstance(op, P.YoloOp):
                yolo_i -= 1
                gv, acc = claim(op.src)
                if acc:
                    raise nat.NativeError("a head convolution feeds more than one consumer")
                self.bwd.append(lambda flat, dps, i=yolo_i, g=gv: T.yolo_train_bwd(dps[i], g))
                continue
            out = getattr(op, "out", None)
            if out is None:
                continue
            gout = grads.get(id(out))
            if gout is None or not is_written(gout):
                continue      # nothing downstream depends on this tensor
            dy = gout.view
            cur["reads"].add(id(root(gout)))
            if isinstance(op, P.ConvOp):
                st = conv_state[id(op)]
                self._conv_bwd(st, dy, claim if not st["stem"] else None)
            elif isinstance(op, P.AddOp):
                m = op.module
                self.aux_bwd.append(("add", op, dy))      # (diagnostics: oracle/layerwise.compare_backward_aux)
                for i, operand in enumerate([op.x, op.others[0]]):
                    gv, acc = claim(operand)
                    self.bwd.append(lambda flat, dps, d=dy, g=gv, a=acc, o=op, i=i:
                                    T.axpby(d, g, o.wall[i:i + 1] if o.wall is not None else None, a))
                if m.weight:
                    self.bwd_writes[len(self.bwd)] = [m.w]
                    self.bwd.append(lambda flat, dps, d=dy, o=op, mm=m:
                                    T.fusion_weights_bwd(d, o.x.view, o.others[0].view, mm.w.detach(), self._pgrad(flat, mm.w)))
            elif isinstance(op, P.ConcatOp):
                off = 0
                for s_ in op.srcs:
                    g = gslot(s_)
                    if g.alias_of is None or g.alias_of is not grads.get(id(op.out)):
                        gv, acc = claim(s_)
                        src = View(dy.buf, dy.c_off + off, s_.C)
                        self.bwd.append(lambda flat, dps, a=src, b=gv, c=acc: T.axpby(a, b, None, c))
                    off += s_.C
            elif isinstance(op, P.PoolOp):
                gv, acc = claim(op.src)
                self.bwd.append(lambda flat, dps, o=op, d=dy, g=gv, a=acc: T.maxpool_bwd(o.src.view, d, g, o.k, o.stride, a))
            elif isinstance(op, P.UpOp):
                gv, acc = claim(op.src)
                self.bwd.append(lambda flat, dps, o=op, d=dy, g=gv, a=acc: T.upsample_bwd(d, g, o.s, a))
            elif isinstance(op, P.SEOp):
                gv, acc = claim(op.src)
                self.aux_bwd.append(("se", op, dy))

                def se(flat, dps, o=op, d=dy, g=gv, a=acc):
                    m = o.module
                    w1, b1, w2, b2 = ops.se_weights(m.fc1, m.fc2)
                    T.se_bwd(o.src.view, d, g, w1, b1, w2, b2, o.pooled, o.gate,
                             self._pgrad(flat, m.fc1.weight).view(w1.shape), self._pgrad(flat, m.fc1.bias),
                             self._pgrad(flat, m.fc2.weight).view(w2.shape), self._pgrad(flat, m.fc2.bias), a)
                self.bwd_writes[len(self.bwd)] = [op.module.fc1.weight, op.module.fc1.bias, op.module.fc2.weight, op.module.fc2.bias]
                self.bwd.append(se)
        close_group()
        self._schedule_lanes_backward()
        self._schedule_ready()

    def _schedule_lanes_backward(self):
        """Cross-lane ordering of the backward groups.  A group runs on the lane of its op.  Every gradient buffer it reads
        (dy) or writes (first writer overwrites, later ones accumulate: the order fixed when the plan was bound must hold)
        may last have been written by a group of the other lane: it then waits for the event recorded after that group."""
        last = {}                        # gradient buffer -> {lane: index of the last group that wrote it}
        self.bwd_waits, self.bwd_records = {}, set()
        for gi, g in enumerate(self.bwd_groups):
            lane = self.lanes[g["op"]]
            for r in g["reads"] | g["writes"]:
                other = last.get(r, {}).get(1 - lane)
                if other is not None:
                    self.bwd_waits.setdefault(gi, set()).add(other)
                    self.bwd_records.add(other)
            for r in g["writes"]:
                last.setdefault(r, {})[lane] = gi
        self._step_group = {}
        for gi, g in enumerate(self.bwd_groups):
            for i in range(g["lo"], g["hi"]):
                self._step_group[i] = gi

    def _schedule_ready(self):
        """ready_after[i] = flat ranges (lo, hi) whose gradients are final once bwd step i has been enqueued; ranges of
        parameters no step writes (their gradient stays zero) are ready from the start (key -1).  Used to start the
        gradient all-reduce of finished buckets while the rest of the backward still runs (dyk/dist_utils.py)."""
        last = {}
        for i, prms in self.bwd_writes.items():
            for prm in prms:
                last[id(prm)] = max(last.get(id(prm), -1), i)
        self.ready_after = {}
        for prm in self.params:
            o = self.offsets[id(prm)]
            n = (prm.numel() + 31) // 32 * 32
            self.ready_after.setdefault(last.get(id(prm), -1), []).append((o, o + n))

    def _conv_bwd(self, st, dy, claim):
        op, conv, bn = st["op"], st["conv"], st["bn"]
        k, s, p = st["k"], st["s"], st["p"]
        gin = claim(op.src) if claim is not None else None
        st["dy_view"], st["gin"] = dy, gin      # kept for the teacher-forced backward check (oracle/layerwise.py)

        def run(flat, dps):
            gw = self._pgrad(flat, conv.weight)
            ring = bn is not None
            if bn is not None:
                z = st["z"]
                dz = self.wg.next_dz(z)
                T.bn_act_bwd(dy, z, st["scale"], st["shift"], st["mean"], st["invstd"], bn.weight.detach(), op.act, dz,
                             self._pgrad(flat, bn.weight), self._pgrad(flat, bn.bias))
                cout_real = None
            else:
                dz = dy                                   # padded 16-bit head gradient
                cout_real = conv.out_channels
                if conv.bias is not None:
                    o = self.offsets[id(conv.bias)]
                    T.chan_sum(dz, flat[o:o + dz.C], accumulate=True)     # dz.C = head_pad(Cout) = the padded slot of the bias
            if st["stem"]:
                T.stem_wgrad_tc(st["x_in"], dz, gw, k=k, stride=s, pad=p, accumulate=True, resize=self.resize)
                return
            gv, acc = gin
            if st["dw"]:
                T.dwconv_wgrad(op.src.view, dz, gw, k=k, stride=s, pad=p, accumulate=True)
                T.dwconv_dgrad(dz, st["w"], gv, k=k, stride=s, pad=p, accumulate=acc)
                return
            self.wg.launch(lambda: T.conv_wgrad(op.src.view, dz, gw, k=k, stride=s, pad=p, accumulate=True, cout_real=cout_real),
                           ring=ring)
            wd = st["wd"] if st.get("multi") else T.pack_dgrad_weight(conv.weight, self.dtype, opad=dz.C)
            T.conv_dgrad(dz, wd, gv, k=k, stride=s, pad=p, accumulate=acc)
        self.bwd_writes[len(self.bwd)] = [q for q in (conv.weight, conv.bias, bn.weight if bn is not None else None,
                                                       bn.bias if bn is not None else None) if q is not None]
        self.bwd.append(run)

    def _backward_range(self, glo, ghi, reducer=None, record_sends=False):
        """Backward groups [glo, ghi) on their lanes; every forked stream has re-joined the main stream on return."""
        main = torch.cuda.current_stream()
        if glo == 0:
            self.flat.zero_()
        if self.two_lanes:
            fork = torch.cuda.Event()
            fork.record(main)
            self.lane1.wait_event(fork)
        streams = (main, self.lane1)
        events = {}
        for gi in range(glo, ghi):
            g = self.bwd_groups[gi]
            lane = self.lanes[g["op"]]
            st = streams[lane]
            for w in self.bwd_waits.get(gi, ()):
                if w >= glo:                       # earlier ranges were joined when they ended
                    st.wait_event(events[w])
            self.wg = self.wgs[lane]
            if lane == 0:
                for i in range(g["lo"], g["hi"]):
                    self.bwd[i](self.flat, self.dps_in)
            else:
                with torch.cuda.stream(st):
                    for i in range(g["lo"], g["hi"]):
                        self.bwd[i](self.flat, self.dps_in)
            if gi in self.bwd_records:
                ev = torch.cuda.Event()
                ev.record(st)
                events[gi] = ev
            if reducer is not None:                # eager pass with a gradient reducer: feed it group by group
                self._join_lanes(main)
                sent = reducer.calls
                for i in range(g["lo"], g["hi"]):
                    reducer.feed(self.ready_after.get(i, ()))
                if record_sends and reducer.calls != sent:
                    self._seg_ends.append(gi)
        self._join_lanes(main)

    def _join_lanes(self, main):
        if self.two_lanes:
            with torch.cuda.stream(self.lane1):
                self.wgs[1].join()                 # lane 1's weight gradients re-join lane 1 ...
            j = torch.cuda.Event()
            j.record(self.lane1)
            main.wait_event(j)                     # ... and lane 1 re-joins the main stream
        self.wgs[0].join()

    def backward(self, dps):
        """The backward launches of one step, as CUDA graph replays.  Without a gradient reducer that is one graph; with
        dyk.dist_utils.OverlappedAllReduce the list is cut after every backward step at which the first (eager) pass saw
        a bucket go out, so the all-reduce of a finished bucket is issued between two replays and overlaps the next one.
        The gradients are returned as views of a COPY of the plan's flat buffer (0.15 ms for 464 MB): the next backward
        overwrites the buffer, and gradient accumulation over several backward passes must not alias it."""
        with T.use_scratch(self._scratch):
            return self._backward(dps)

    def _backward(self, dps):
        for d, buf in zip(dps, self.dps_in):
            if d is None:
                buf.zero_()
            else:
                buf.copy_(d.detach())
        flat, n = self.flat, len(self.bwd)
        reducer = self.reducer
        active = reducer is not None and reducer._active()
        key = (id(reducer), getattr(reducer, "bucket_bytes", 0), active)
        if key != self._bwd_key:
            self._bwd_key, self._bwd_graphs, self._bwd_warm = key, None, False
        if reducer is not None:
            reducer.begin(flat, self.grad_numel)
            reducer.feed(self.ready_after.get(-1, ()))
        ng = len(self.bwd_groups)
        if self._graphs_on() and self._bwd_warm:
            if self._bwd_graphs is None:
                cuts = [0] + [e + 1 for e in self._seg_ends if e + 1 < ng] + [ng]
                graphs = []
                for lo, hi in zip(cuts, cuts[1:]):
                    g = self._capture(lambda lo=lo, hi=hi: self._backward_range(lo, hi))
                    if g is None:
                        break
                    graphs.append((lo, hi, g))
                self._bwd_graphs = graphs if len(graphs) == len(cuts) - 1 else None
        if self._graphs_on() and self._bwd_warm and self._bwd_graphs is not None:
            for lo, hi, g in self._bwd_graphs:
                g.replay()
                if reducer is not None:
                    for gi in range(lo, hi):
                        for i in range(self.bwd_groups[gi]["lo"], self.bwd_groups[gi]["hi"]):
                            reducer.feed(self.ready_after.get(i, ()))
            nat.count_launches(self._bwd_launches)
        else:
            n0 = nat.launch_count()
            self._seg_ends = []
            self._backward_range(0, ng, reducer=reducer, record_sends=True)
            self._bwd_launches, self._bwd_warm = nat.launch_count() - n0, True
        if reducer is not None:
            reducer.finish()
        out = flat.clone()
        self.last_flat = out
        return [self._pgrad(out, prm) if prm.requires_grad else None for prm in self.params]


class _TrainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, x, y, *params):
        ctx.plan = plan
        ctx.n_params = len(params)
        outs = plan.forward(x, y)
        return outs

    @staticmethod
    def backward(ctx, *dps):
        with torch.cuda.device(ctx.plan.device):
            grads = ctx.plan.backward(list(dps))
        return (None, None, None, *grads)


class TrainPlanCache:
    """At most `keep` training plans are alive (each holds every activation of a step); multi-scale training
    (reference train.py:143-151) switches shapes every few iterations, so the least recently used plan is dropped."""

    def __init__(self, model, keep=2):
        self.model, self.keep = model, keep
        self.plans = {}
        self.last_plan = None

    def invalidate(self):
        self.plans.clear()
        self.last_plan = None

    def run(self, x, y, input_size=None):
        model = self.model
        ops._require_cuda(x, "YOLO.forward (training)")
        dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else model.compute_dtype
        if dtype not in (torch.float16, torch.bfloat16):
            raise nat.NativeError(f"compute dtype {dtype} is not supported (float16 / bfloat16)")
        if y is not None and (y.shape != x.shape or y.device != x.device):
            raise ValueError("visible and LWIR batches must have the same shape and device")

        def prep(t):
            if t is None:
                return None
            if t.dtype not in (torch.float32, torch.uint8):
                t = t.float()
            return t.detach().contiguous()

        x, y = prep(x), prep(y)
        B, _, Hs, Ws = x.shape
        H, W = Hs, Ws
        if input_size is not None:
            H, W = (int(input_size), int(input_size)) if isinstance(input_size, int) else (int(input_size[0]), int(input_size[1]))
        if y is not None and y.dtype != x.dtype:
            raise ValueError("visible and LWIR batches must have the same dtype")
        key = (B, H, W, dtype, y is not None, x.device, x.dtype, Hs, Ws)
        plan = self.plans.pop(key, None)
        with torch.cuda.device(x.device):
            if plan is None:
                while len(self.plans) >= self.keep:
                    self.plans.pop(next(iter(self.plans)))
                plan = TrainPlan(model, B, H, W, dtype, y is not None, x.device, in_dtype=x.dtype, src_hw=(Hs, Ws))
            self.plans[key] = plan           # re-insert = most recently used
            self.last_plan = plan
            plan.reducer = getattr(model, "grad_reducer", None)
            outs = _TrainFunction.apply(plan, x, y, *plan.params)
        return list(outs)
