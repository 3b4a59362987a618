"""Execution-plan compiler for ``models.YOLO``.

The reference executes a model by walking ``module_list`` and keeping a side table ``out[]`` of routed
tensors (models.py:279-315), one library kernel per primitive.  Here the same block list is compiled,
per (mode, batch, H, W, dtype), into a short list of native launches over channels-last buffers:

  * [convolutional] = conv + folded BN + activation in one tcgen05 kernel; a following un-weighted
    [shortcut] becomes the conv's residual operand and a following [upsample] x2 its store pattern;
  * [route] with one source is an alias; with several sources the *producers* are redirected to write
    straight into channel slices of the concat buffer, so no concat kernel runs at all (a copy is only
    emitted when a tensor would have to live in two concat buffers);
  * weighted [shortcut] (modality fusion), [maxpool], [se], [depthwiseconvolutional], [yolo] each map to
    their kernel; the three heads write into one (B, rows, no) prediction tensor;
  * buffers are recycled by liveness, and everything after the two stem convolutions is captured in a
    CUDA graph (the stems read the caller's NCHW tensors directly and therefore stay outside).

Parameters stay in the stock nn modules; packed / BN-folded copies live in a WeightBank that is refreshed
in place whenever a parameter's version counter or storage changes.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn

from . import _native as nat
from . import ops
from .ops import View


class Value:
    """One logical activation tensor (N, H, W, C) produced by one op."""
    __slots__ = ("C", "H", "W", "uses", "ext", "f32", "place", "view", "kind", "first", "last", "name", "virtual")

    def __init__(self, C_, H, W, kind, name="", ext=False, f32=False):
        self.C, self.H, self.W = C_, H, W
        self.uses = 0
        self.ext = ext
        self.f32 = f32
        self.place = None   # (parent Value, channel offset) when it lives inside a concat buffer
        self.view = None
        self.kind = kind
        self.first = None
        self.last = None
        self.name = name
        self.virtual = False   # never materialised: formed (and rounded to the storage type) inside its consumer


class Op:
    kind = "op"

    def inputs(self):
        return []


class ConvOp(Op):
    kind = "conv"

    def __init__(self, layer, src, out, conv, bn, act, tag=""):
        self.layer, self.src, self.out, self.conv, self.bn, self.act = layer, src, out, conv, bn, act
        self.res = None
        self.upsample2x = False
        self.out_f32 = False
        self.tag = tag
        self.src2 = None        # dual-source input: src = w0 * src + w1 * src2 (fused WeightedFeatureFusion)
        self.fusion = None      # ... and the module that owns the raw weights
        self.gate_of = None     # SEOp whose gate multiplies src on the fly (fused SqueezeExcitation scale)

    @property
    def flavor(self):
        c = self.conv
        if c.groups == 1:
            return "stem" if self.src.ext else "dense"
        if c.groups == c.in_channels == c.out_channels:
            return "dw"
        raise nat.NativeError(f"layer {self.layer}: grouped convolution groups={c.groups} "
                              f"({c.in_channels}->{c.out_channels}) has no native kernel")

    def inputs(self):
        return [self.src] + ([self.src2] if self.src2 is not None else []) + ([self.res] if self.res is not None else []) \
            + ([self.gate_of.out] if self.gate_of is not None else [])      # (the gate: a virtual Value the SE op produces)


class AddOp(Op):
    kind = "add"

    def __init__(self, layer, x, others, out, module):
        self.layer, self.x, self.others, self.out, self.module = layer, x, others, out, module
        self.gate_of = None     # SEOp whose gate multiplies x on the fly

    def inputs(self):
        return [self.x] + list(self.others) + ([self.gate_of.out] if self.gate_of is not None else [])


class ConcatOp(Op):
    kind = "concat"

    def __init__(self, layer, srcs, out):
        self.layer, self.srcs, self.out = layer, srcs, out
        self.copies = []  # (src Value, channel offset) that could not be placed

    def inputs(self):
        return list(self.srcs)


class PoolOp(Op):
    kind = "maxpool"

    def __init__(self, layer, src, out, k, stride):
        self.layer, self.src, self.out, self.k, self.stride = layer, src, out, k, stride

    def inputs(self):
        return [self.src]


class UpOp(Op):
    kind = "upsample"

    def __init__(self, layer, src, out, s):
        self.layer, self.src, self.out, self.s = layer, src, out, s

    def inputs(self):
        return [self.src]


class SEOp(Op):
    kind = "se"

    def __init__(self, layer, src, out, module):
        self.layer, self.src, self.out, self.module = layer, src, out, module
        self.gate_only = False  # True: only gate[n, c] is computed; every consumer applies it while loading (fuse_se_gates)
        self.gate = None

    def inputs(self):
        return [self.src]


class YoloOp(Op):
    kind = "yolo"

    def __init__(self, layer, src, module, index):
        self.layer, self.src, self.module, self.index = layer, src, module, index
        self.out = None

    def inputs(self):
        return [self.src]


def _conv_out(h, k, s, p):
    return (h + 2 * p - k) // s + 1


# ---------------------------------------------------------------------------------------------- graph building
def build_ops(model, H, W, dual):
    """Block list -> raw op list + the Value visible at every layer index (the reference's `out[]`)."""
    defs, mods = model.module_defs, model.module_list
    second = model.net_info.get("second_index") if dual else None
    if not dual and "second_index" in model.net_info:
        # the reference would feed a deep feature map into the 3-channel LWIR stem and fail in conv2d
        raise ValueError("this cfg defines second_index: call model(visible, lwir) with both modalities")
    img0 = Value(3, H, W, "input", "visible", ext=True)
    img1 = Value(3, H, W, "input", "lwir", ext=True) if dual else None
    ops_, vals = [], []
    cur = img0
    yolo_idx = 0

    def use(v):
        v.uses += 1
        return v

    def add_conv(layer, src, conv, bn, act, tag=""):
        k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        if conv.kernel_size[0] != conv.kernel_size[1] or conv.stride[0] != conv.stride[1]:
            raise nat.NativeError(f"layer {layer}: non-square kernels/strides have no native kernel")
        if conv.in_channels != src.C:
            raise ValueError(f"layer {layer}: conv expects {conv.in_channels} channels, got {src.C}")
        out = Value(conv.out_channels, _conv_out(src.H, k, s, p), _conv_out(src.W, k, s, p), "conv", f"L{layer}{tag}")
        ops_.append(ConvOp(layer, use(src), out, conv, bn, act, tag))
        return out

    def add_cba(layer, src, cba, tag):
        mods_ = list(cba.conv)
        bn = mods_[1] if len(mods_) > 1 and isinstance(mods_[1], nn.BatchNorm2d) else None
        from build_utils.layers import activation_name
        return add_conv(layer, src, mods_[0], bn, activation_name(mods_[-1]), tag)

    for i, (d, m) in enumerate(zip(defs, mods)):
        t = d["type"]
        if t == "convolutional":
            src = img1 if (second is not None and i == second) else cur
            bn = m[1] if d["batch_normalize"] else None
            cur = add_conv(i, src, m[0], bn, d["activation"])
        elif t == "depthwiseconvolutional":
            mid = add_conv(i, cur, m.conv[0], m.conv[1], "relu6", ".dw")
            cur = add_conv(i, mid, m.conv[3], m.conv[4], "relu6", ".pw")
        elif t == "inception":
            x = cur
            b1 = add_cba(i, x, m.branch1[0], ".b1")
            b2 = add_cba(i, add_cba(i, x, m.branch2[0], ".b2a"), m.branch2[1], ".b2b")
            b3 = add_cba(i, add_cba(i, add_cba(i, x, m.branch3[0], ".b3a"), m.branch3[1], ".b3b"), m.branch3[2], ".b3c")
            pooled = Value(x.C, x.H, x.W, "maxpool", f"L{i}.pool")
            ops_.append(PoolOp(i, use(x), pooled, 3, 1))
            b4 = add_cba(i, pooled, m.branch4[1], ".b4")
            srcs = [b1, b2, b3, b4]
            cur = Value(sum(s.C for s in srcs), x.H, x.W, "concat", f"L{i}")
            ops_.append(ConcatOp(i, [use(s) for s in srcs], cur))
        elif t == "dropout":
            if model.training and getattr(m, "p", 0.0) > 0:
                raise nat.NativeError(f"layer {i}: [dropout] with p = {m.p} has no native training kernel (eval mode: identity)")
        elif t == "se":
            out = Value(cur.C, cur.H, cur.W, "se", f"L{i}")
            ops_.append(SEOp(i, use(cur), out, m))
            cur = out
        elif t == "maxpool":
            k, s = d["size"], d["stride"]
            p = (k - 1) // 2
            out = Value(cur.C, _conv_out(cur.H, k, s, p), _conv_out(cur.W, k, s, p), "maxpool", f"L{i}")
            ops_.append(PoolOp(i, use(cur), out, k, s))
            cur = out
        elif t == "upsample":
            s = d["stride"]
            out = Value(cur.C, cur.H * s, cur.W * s, "upsample", f"L{i}")
            ops_.append(UpOp(i, use(cur), out, s))
            cur = out
        elif t == "route":
            srcs = [vals[l] for l in m.layers]
            if len(srcs) == 1:
                cur = srcs[0]  # alias, like FeatureConcat with a single layer (layers.py:44)
            else:
                if any((s.H, s.W) != (srcs[0].H, srcs[0].W) for s in srcs):
                    raise ValueError(f"layer {i}: route sources have different spatial sizes")
                cur = Value(sum(s.C for s in srcs), srcs[0].H, srcs[0].W, "concat", f"L{i}")
                ops_.append(ConcatOp(i, [use(s) for s in srcs], cur))
        elif t == "shortcut":
            others = [vals[l] for l in m.layers]
            out = Value(cur.C, cur.H, cur.W, "add", f"L{i}")
            ops_.append(AddOp(i, use(cur), [use(o) for o in others], out, m))
            cur = out
        elif t == "yolo":
            ops_.append(YoloOp(i, use(cur), m, yolo_idx))
            yolo_idx += 1
        elif t == "avgpool":
            raise nat.NativeError(f"layer {i}: [avgpool] is not part of any detection cfg and has no native kernel")
        else:
            pass  # the reference only prints a warning and passes x through an empty Sequential
        vals.append(cur)
    return ops_, vals, img0, img1


def fuse(ops_):
    """Peephole fusion: conv -> (+ residual) -> (upsample x2)."""
    out, k = [], 0
    while k < len(ops_):
        op = ops_[k]
        if isinstance(op, ConvOp) and op.flavor == "dense":
            nxt = ops_[k + 1] if k + 1 < len(ops_) else None
            if (isinstance(nxt, AddOp) and not nxt.module.weight and len(nxt.others) == 1 and op.out.uses == 1
                    and (nxt.x is op.out or nxt.others[0] is op.out) and nxt.x is not nxt.others[0]):
                other = nxt.others[0] if nxt.x is op.out else nxt.x
                if (other.C, other.H, other.W) == (op.out.C, op.out.H, op.out.W) and not other.ext:
                    op.res = other
                    op.out = nxt.out
                    op.out.kind = "conv"
                    k += 1
                    nxt = ops_[k + 1] if k + 1 < len(ops_) else None
            if (isinstance(nxt, UpOp) and nxt.s == 2 and nxt.src is op.out and op.out.uses == 1
                    and op.res is None):
                op.upsample2x = True
                op.out = nxt.out
                op.out.kind = "conv"
                k += 1
        out.append(op)
        k += 1
    return out


def fuse_weighted_adds(ops_, B, dtype):
    """Weighted [shortcut] (the modality fusion `x * w[0] + a * w[1]`, build_utils/layers.py:63-85) whose only consumer is
    a convolution with a dual-source operand path (dyk_conv2d_dual_source_supported): the sum is formed inside that
    convolution while its operand is staged, so the add launch, the write of the sum and its re-read all disappear."""
    if os.environ.get("DYK_FUSE_ADD", "1") == "0":
        return ops_
    consumers = {}
    for op in ops_:
        for v in op.inputs():
            consumers.setdefault(id(v), []).append(op)
    drop = set()
    for op in ops_:
        if not (isinstance(op, AddOp) and op.module.weight and len(op.others) == 1):
            continue
        a, b, out = op.x, op.others[0], op.out
        cons = consumers.get(id(out), [])
        if len(cons) != 1 or out.uses != 1 or a.ext or b.ext or a.f32 or b.f32:
            continue
        c = cons[0]
        if not (isinstance(c, ConvOp) and c.flavor == "dense" and c.src is out and c.res is None and c.src2 is None):
            continue
        if (a.C, a.H, a.W) != (b.C, b.H, b.W) or a.C != out.C:
            continue
        k, s, p = c.conv.kernel_size[0], c.conv.stride[0], c.conv.padding[0]
        if not ops.conv_dual_source_supported(B, a.H, a.W, a.C, c.conv.out_channels, c.out.C, dtype, k=k, stride=s, pad=p,
                                              upsample2x=c.upsample2x, out_f32=c.out_f32):
            continue
        c.src, c.src2, c.fusion = a, b, op.module
        out.virtual = True
        drop.add(id(op))
    return [op for op in ops_ if id(op) not in drop]


def fuse_se_gates(ops_, B, dtype):
    """SqueezeExcitation (layers.py:184-190) without its `scale * x` pass: when every consumer of an [se] block can take the
    gate itself — a 1x1 convolution (conv(x * g_n) = x . (W * g_n): per-image weights, dyk_conv_params.w_image_stride) or
    a [shortcut] whose first operand is the block's output (dyk_fused_add_gated) — the block only computes gate[n, c] and
    the consumers read the block's INPUT.  The gated activation tensor (the largest tensor of the block) is never written
    or re-read.  Rounding: a gated [shortcut] rounds x * g exactly like the materialised form; a gated convolution rounds
    W * g_n instead of x * g_n (same error class; the oracle models it, oracle/darknet_ref.py `gated`)."""
    if os.environ.get("DYK_FUSE_SE", "1") == "0":
        return ops_
    consumers = {}
    for op in ops_:
        for v in op.inputs():
            consumers.setdefault(id(v), []).append(op)
    for se in ops_:
        if not isinstance(se, SEOp) or se.src.ext or se.src.f32:
            continue
        out, x = se.out, se.src
        cons = consumers.get(id(out), [])
        if not cons or out.uses != len(cons):
            continue
        ok = True
        for c in cons:
            if isinstance(c, ConvOp) and c.flavor == "dense" and c.src is out and c.res is not out and c.src2 is None \
                    and c.gate_of is None:
                k, s, p = c.conv.kernel_size[0], c.conv.stride[0], c.conv.padding[0]
                ok = ok and ops.conv_gated_input_supported(x.H, x.W, x.C, k=k, stride=s, pad=p, upsample2x=c.upsample2x,
                                                           out_f32=c.out_f32)
            elif isinstance(c, AddOp) and c.x is out and len(c.others) == 1 and c.others[0] is not out and c.gate_of is None \
                    and c.others[0].C == out.C:
                pass
            else:
                ok = False
            if not ok:
                break
        if not ok:
            continue
        for c in cons:
            if isinstance(c, ConvOp):
                c.src = x
            else:
                c.x = x
            c.gate_of = se
            x.uses += 1
        out.virtual = True
        se.gate_only = True
    return ops_


def cascade_pools(ops_):
    """SPP (models.py:91-94): the 5x5, 9x9 and 13x13 stride-1 max pools of one tensor.  Max pooling with -inf padding
    composes — mp5(mp5(x)) = mp9(x), mp5(mp9(x)) = mp13(x), bit for bit — so the larger windows are computed from the
    previous pool's output with a 5x5 window: 25 loads per output instead of 81 / 169."""
    by_src = {}
    for op in ops_:
        if isinstance(op, PoolOp) and op.stride == 1:
            prev = by_src.get((id(op.src), op.k - 4))
            by_src[(id(op.src), op.k)] = op
            if prev is not None and op.k in (9, 13):
                op.src.uses -= 1
                op.src = prev.out
                op.src.uses += 1
                op.k = 5


def mark_heads(ops_):
    """Head convs (no BN, linear) that only feed [yolo] blocks keep their logits in fp32."""
    consumers = {}
    for op in ops_:
        for v in op.inputs():
            consumers.setdefault(id(v), []).append(op)
    for op in ops_:
        if isinstance(op, ConvOp) and op.flavor == "dense" and op.bn is None and op.act == "linear" \
                and op.res is None and not op.upsample2x:
            cons = consumers.get(id(op.out), [])
            if cons and all(isinstance(c, YoloOp) for c in cons):
                op.out_f32 = True
                op.out.f32 = True


def place_concats(ops_):
    """Let producers write directly into the channel slices of concat buffers."""
    for op in ops_:
        if not isinstance(op, ConcatOp):
            continue
        off = 0
        seen = set()
        for s in op.srcs:
            can_place = (s.place is None and not s.ext and s.kind != "concat" and not s.f32 and id(s) not in seen
                         and s.C % 8 == 0 and off % 8 == 0)
            if can_place:
                s.place = (op.out, off)
            else:
                op.copies.append((s, off))
            seen.add(id(s))
            off += s.C


def liveness(ops_):
    for k, op in enumerate(ops_):
        outs = [op.out] if getattr(op, "out", None) is not None else []
        for v in outs:
            if v.first is None:
                # stem outputs are written by the eager stem launches that precede the (graph) body
                v.first = -1 if (isinstance(op, ConvOp) and op.flavor == "stem") else k
            v.last = k if v.last is None else max(v.last, k)
        for v in op.inputs():
            v.last = k if v.last is None else max(v.last, k)
            if v.first is None:
                v.first = 0


def assign_lanes(ops_):
    """Two-lane schedule of the op list.  The dual-stream models are two independent backbones between fusion points
    (SURVEY.md Appendix A), and at batch 16 the late layers have too few tiles to fill 148 SMs for whole waves; running
    the visible and the LWIR branch on two CUDA streams lets the CTAs of one branch fill the tail of the other.

    Greedy rule over the (already topologically ordered) ops, using transitive-ancestor bitsets: an op goes to the lane
    whose most recent op it depends on; if it depends on both it joins on lane 0; if on neither it stays with its most
    recent ancestor (short CSP side branches), and an op without any ancestor in the body (the first layer after the
    LWIR stem) opens lane 1.  Returns (lane per op, producer-op index per Value id)."""
    producer = {}
    anc = []
    lanes = []
    last = [None, None]
    for k, op in enumerate(ops_):
        a = 0
        for v in op.inputs():
            p = producer.get(id(v))
            if p is not None and lanes[p] >= 0:           # stems are sources: they run before the body
                a |= (1 << p) | anc[p]
        anc.append(a)
        out = getattr(op, "out", None)
        if out is not None:
            producer[id(out)] = k
        if isinstance(op, ConvOp) and op.flavor == "stem":
            lanes.append(-1)                      # runs eagerly before the body, on the caller's stream
            continue
        d0 = last[0] is not None and (a >> last[0]) & 1
        d1 = last[1] is not None and (a >> last[1]) & 1
        if d0 and not d1:
            lane = 0
        elif d1 and not d0:
            lane = 1
        elif d0 and d1:
            lane = 0
        elif a:
            lane = max(lanes[a.bit_length() - 1], 0)      # most recent ancestor's lane
        else:
            lane = 0 if last[0] is None else 1
        lanes.append(lane)
        last[lane] = k
    return lanes, producer, anc


def dual_phase(lanes, anc):
    """For every op: is there an op on the OTHER lane that neither depends on it nor is depended on by it, i.e. can the two
    lanes be busy at the same time here?  (True for the two modality backbones between fusion points, False for the neck /
    heads.)  Convolutions of such ops are launched with half the SMs each (dyk_conv_params.sm_limit): both lanes' kernels
    are then resident together, a kernel boundary on one lane idles half the machine instead of all of it, and the tiles
    of the two launches are rounded up over 74 SMs each instead of one after the other over 148."""
    n = len(lanes)
    by_lane = {0: [k for k in range(n) if lanes[k] == 0], 1: [k for k in range(n) if lanes[k] == 1]}
    out = [False] * n
    for k in range(n):
        if lanes[k] < 0:
            continue
        for j in by_lane[1 - lanes[k]]:
            if not (anc[k] >> j) & 1 and not (anc[j] >> k) & 1:
                out[k] = True
                break
    return out


class WeightBank:
    """Packed weights / folded BN vectors per conv module, refreshed in place when parameters change."""

    def __init__(self):
        self.entries = {}     # (id(conv), dtype, flavor) -> dict
        self.pending = []     # entries allocated by get() and not yet filled
        self.signature = None

    @staticmethod
    def model_signature(tensors):
        return tuple(t._version for t in tensors)

    def get(self, conv, bn, dtype, flavor):
        """Entry with its device buffers allocated; the contents are written by the next flush()."""
        key = (id(conv), dtype, flavor)
        e = self.entries.get(key)
        if e is None:
            e = {"conv": conv, "bn": bn, "dtype": dtype, "flavor": flavor}
            k = conv.kernel_size[0]
            dev = conv.weight.device
            wt = conv.weight
            if flavor == "dense" and dtype != torch.float32 and k * k <= 9 and wt.dtype == torch.float32 and wt.is_contiguous():
                e["w"] = torch.empty((conv.out_channels, k, k, conv.in_channels), dtype=dtype, device=dev)
                e["multi"] = True
            else:
                e["w"] = self._pack_single(e)
                e["multi"] = False
            e["scale"], e["bias"] = ops.fold_bn_alloc(conv, bn)
            self.entries[key] = e
            self.pending.append(e)
        return e

    @staticmethod
    def _pack_single(e):
        conv, dtype, flavor = e["conv"], e["dtype"], e["flavor"]
        k = conv.kernel_size[0]
        if flavor == "dense":
            return ops.f32_pack_conv_weight(conv.weight) if dtype == torch.float32 else ops.pack_conv_weight(conv.weight, dtype)
        if flavor == "stem":
            return conv.weight.detach().float().permute(0, 2, 3, 1).contiguous()
        # depthwise: [k][k][C] fp32
        return conv.weight.detach().float().reshape(conv.out_channels, k, k).permute(1, 2, 0).contiguous()

    def _fill(self, entries, first):
        """Two native launches for the whole list: every dense 16-bit weight re-laid out (dyk_pack_weights_multi) and
        every BatchNorm folded (dyk_fold_bn_multi); the few stem / depthwise / fp32-mode weights are permuted one by one.
        Device addresses never change: captured graphs hold them."""
        by_dtype = {}
        for e in entries:
            if e["multi"]:
                by_dtype.setdefault(e["dtype"], []).append((e["conv"].weight.detach(), e["w"]))
            elif not first:
                e["w"].copy_(self._pack_single(e))
        for dtype, items in by_dtype.items():
            ops.pack_conv_weights_multi(items, dtype)
        native = []
        for e in entries:
            conv, bn = e["conv"], e["bn"]
            ts = [bn.weight, bn.bias, bn.running_mean, bn.running_var] if bn is not None else [conv.bias]
            if all(t is None or (t.dtype == torch.float32 and t.is_contiguous()) for t in ts):
                native.append((conv, bn, e["scale"], e["bias"]))
            else:                                   # .half() models: fold with torch ops in fp32
                scale, bias = ops.fold_bn(conv, bn)
                if scale is not None:
                    e["scale"].copy_(scale)
                if bias is not None:
                    e["bias"].copy_(bias)
        ops.fold_bn_multi(native)

    def flush(self):
        if self.pending:
            self._fill(self.pending, first=True)
            self.pending = []

    def refresh(self):
        self.pending = []
        self._fill(list(self.entries.values()), first=False)


class Plan:
    def __init__(self, model, bank, B, H, W, dtype, dual, device):
        self.B, self.dtype, self.device, self.dual = B, dtype, device, dual
        self.H, self.W = H, W
        raw, vals, self.img0, self.img1 = build_ops(model, H, W, dual)
        self.layer_vals = vals   # Value visible at every cfg layer index (diagnostics: tools/layer_parity.py)
        # fp32-accurate mode (compute_dtype = torch.float32, csrc/f32_path.cu): fp32 buffers, split-bf16 tensor-core convs,
        # no residual / upsample fusion into the conv epilogue (its fp32 store path is the plain one), one lane
        self.f32 = dtype == torch.float32
        self.ops = raw if self.f32 else fuse_se_gates(fuse_weighted_adds(fuse(raw), B, dtype), B, dtype)
        cascade_pools(self.ops)
        mark_heads(self.ops)
        place_concats(self.ops)
        liveness(self.ops)
        self.two_lanes = os.environ.get("DYK_LANES", "2") != "1" and not self.f32
        self.lanes, self.producer, anc = assign_lanes(self.ops)
        if not self.two_lanes:
            self.lanes = [min(l, 0) for l in self.lanes]
        # measured (A/B in one session, bs 16 / 64): splitting the machine is 3 - 5 % SLOWER on all three dual models (dyolov3
        # 3110 -> 3015 frames/s, dyolov4 2363 -> 2238, MobileNetV3 6292 -> 6019): the lanes drift, and a lane that is in a
        # light layer or at a kernel boundary lends its SMs to the other one only when grids are not capped.  Off by default.
        split = self.two_lanes and device.type == "cuda" and os.environ.get("DYK_SM_SPLIT", "0") == "1"
        self.dual = dual_phase(self.lanes, anc) if split else [False] * len(self.ops)
        self.half_sms = (torch.cuda.get_device_properties(device).multi_processor_count // 2) if split else 0
        self._allocate()
        self._bind(model, bank)
        bank.flush()
        self.graph = None
        self.graph_failed = False

    # ---- storage --------------------------------------------------------------------------------
    def _allocate(self):
        reuse = os.environ.get("DYK_NO_REUSE", "0") != "1"
        roots = {}
        order = []
        for k, op in enumerate(self.ops):
            v = getattr(op, "out", None)
            if v is None or v.virtual:          # a virtual value (gate-only SE output) owns no buffer
                continue
            root = v.place[0] if v.place else v
            if id(root) not in roots:
                roots[id(root)] = [root, v.first, v.last, set()]
                order.append(id(root))
            r = roots[id(root)]
            r[1] = min(r[1], v.first)
            r[2] = max(r[2], v.last if v.last is not None else v.first)
        # lanes that touch each root (writers and readers); stems (-1) run before the body on the caller's stream
        def root_of(v):
            return v.place[0] if v.place else v
        for k, op in enumerate(self.ops):
            lane = self.lanes[k]
            for v in ([getattr(op, "out", None)] if getattr(op, "out", None) is not None else []) + list(op.inputs()):
                r = roots.get(id(root_of(v)))
                if r is not None and lane >= 0:
                    r[3].add(lane)
        # a concat root's own first/last (its ConcatOp index and its consumers) were merged above through
        # the ConcatOp's `out`; members extend the interval backwards to their producers.
        # Buffers are recycled only within one lane (stream order then guarantees that every access of the previous
        # owner has completed); tensors touched by both lanes are never recycled.
        free = {}     # (shape, dtype, lane) -> list of (tensor, free_after_op)
        self.bytes_allocated = 0
        for rid in sorted(order, key=lambda r: roots[r][1]):
            root, first, last, lanes = roots[rid]
            dt = torch.float32 if (root.f32 or self.f32) else self.dtype
            shape = (self.B, root.H, root.W, root.C)
            lane = next(iter(lanes)) if len(lanes) == 1 else None
            buf = None
            if reuse and lane is not None:
                lst = free.get((shape, dt, lane), [])
                for j, (t, free_after) in enumerate(lst):
                    if free_after < first:
                        buf = t
                        lst.pop(j)
                        break
            if buf is None:
                buf = torch.empty(shape, dtype=dt, device=self.device)
                self.bytes_allocated += buf.numel() * buf.element_size()
            if lane is not None:
                free.setdefault((shape, dt, lane), []).append((buf, last))
            root.view = View(buf, 0, root.C)
        for op in self.ops:
            v = getattr(op, "out", None)
            if v is not None and v.place:
                parent, off = v.place
                v.view = View(parent.view.buf, off, v.C)

    # ---- launches -------------------------------------------------------------------------------
    def _bind(self, model, bank):
        self.stem_steps, self.steps = [], []
        if self.f32:      # one scratch for the split operand [x1|x2|x3|x1|x2|x1] of the widest dense convolution input
            need = max((self.B * op.src.H * op.src.W * 6 * op.src.C for op in self.ops
                        if isinstance(op, ConvOp) and op.flavor == "dense"), default=0)
            self.split_buf = torch.empty(max(need, 1), dtype=torch.bfloat16, device=self.device)
            self.bytes_allocated += need * 2
        self.yolo = []
        yolos = [op for op in self.ops if isinstance(op, YoloOp)]
        rows_total = sum(op.module.na * op.src.H * op.src.W for op in yolos)
        no = yolos[0].module.no if yolos else 0
        self.io = torch.empty((self.B, rows_total, no), dtype=torch.float32, device=self.device) if yolos else None
        self.p_outs = []
        row_off = 0
        dev = self.device
        self.side_stream = torch.cuda.Stream(device=dev) if (self.two_lanes and dev.type == "cuda" and 1 in self.lanes) else None
        events = {}           # producer op index -> event recorded right after its last launch
        synced = {0: -1, 1: -1}   # per consuming lane: newest producer index of the other lane already waited for
        for k, op in enumerate(self.ops):
            lane = self.lanes[k]
            n_before = len(self.steps)
            if lane >= 0 and self.side_stream is not None:
                need = -1
                for v in op.inputs():
                    p = self.producer.get(id(v))
                    if p is not None and self.lanes[p] >= 0 and self.lanes[p] != lane:
                        need = max(need, p)
                if need > synced[lane]:
                    # wait for everything the other lane has launched up to (and including) op `need`
                    ev = events.get(need)
                    if ev is None:
                        raise nat.NativeError("internal: cross-lane producer without an event")
                    self.steps.append(_Sync("wait", ev, lane))
                    synced[lane] = need
            self._bind_op(op, k, bank, rows_total, row_off_ref := [row_off])
            row_off = row_off_ref[0]
            for st in self.steps[n_before:]:
                if not isinstance(st, _Sync):
                    st.lane = max(lane, 0)
                    if isinstance(st, _ConvStep) and self.dual[k]:
                        st.kw["sm_limit"] = self.half_sms
            if lane >= 0 and self.side_stream is not None and self._has_cross_lane_consumer(k):
                ev = torch.cuda.Event()
                events[k] = ev
                self.steps.append(_Sync("record", ev, lane))
        self.launches_per_forward = len(self.stem_steps) + sum(s.launches for s in self.steps)

    def _has_cross_lane_consumer(self, k):
        if not hasattr(self, "_cross"):
            self._cross = set()
            for j, op in enumerate(self.ops):
                for v in op.inputs():
                    p = self.producer.get(id(v))
                    if p is not None and self.lanes[p] >= 0 and self.lanes[j] >= 0 and self.lanes[p] != self.lanes[j]:
                        self._cross.add(p)
        return k in self._cross

    def _bind_op(self, op, k, bank, rows_total, row_off_ref):
        dev = self.device
        row_off = row_off_ref[0]
        if True:
            if isinstance(op, ConvOp):
                fl = op.flavor
                e = bank.get(op.conv, op.bn, self.dtype, fl)
                k, s, p = op.conv.kernel_size[0], op.conv.stride[0], op.conv.padding[0]
                if fl == "stem":
                    which = 0 if op.src is self.img0 else 1
                    self.stem_steps.append((which, e, op.out.view, dict(k=k, stride=s, pad=p, act=op.act)))
                elif self.f32 and fl == "dense":
                    self.steps.append(_Call(ops.f32_conv, op.src.view, e["w"], e["scale"], e["bias"], op.out.view, self.split_buf,
                                            k=k, stride=s, pad=p, act=op.act, cout=op.conv.out_channels))
                    self.steps[-1].launches = 2
                elif self.f32:
                    self.steps.append(_Call(ops.f32_dwconv, op.src.view, e["w"], e["scale"], e["bias"], op.out.view,
                                            k=k, stride=s, pad=p, act=op.act))
                elif fl == "dense":
                    kw = dict(k=k, stride=s, pad=p, act=op.act, res=op.res.view if op.res is not None else None,
                              upsample2x=op.upsample2x, out_f32=op.out_f32)
                    if op.src2 is not None:
                        kw.update(x2=op.src2.view, x_wts_raw=op.fusion.w)
                    if op.gate_of is not None:       # SE gate folded into per-image weights (1x1): scale kernel + conv
                        cout, cin = op.conv.out_channels, op.conv.in_channels
                        wimg = torch.empty((self.B, cout, 1, 1, cin), dtype=self.dtype, device=dev)
                        self.steps.append(_GatedConvStep(op.src.view, e, op.out.view, kw, op.gate_of, wimg))
                    else:
                        self.steps.append(_ConvStep(op.src.view, e, op.out.view, kw))
                else:
                    self.steps.append(_Call(ops.nhwc_dwconv, op.src.view, e["w"], e["scale"], e["bias"], op.out.view,
                                            k=k, stride=s, pad=p, act=op.act))
            elif isinstance(op, AddOp):
                self._bind_add(op)
            elif isinstance(op, ConcatOp):
                for s, off in op.copies:
                    if s.f32:
                        raise nat.NativeError(f"layer {op.layer}: cannot concatenate an fp32 head tensor")
                    self.steps.append(_Call(ops.f32_copy if self.f32 else ops.nhwc_copy, s.view,
                                            View(op.out.view.buf, op.out.view.c_off + off, s.C)))
            elif isinstance(op, PoolOp):
                self.steps.append(_Call(ops.f32_maxpool if self.f32 else ops.nhwc_maxpool, op.src.view, op.out.view, op.k, op.stride))
            elif isinstance(op, UpOp):
                self.steps.append(_Call(ops.f32_upsample if self.f32 else ops.nhwc_upsample, op.src.view, op.out.view, op.s))
            elif isinstance(op, SEOp):
                w1, b1, w2, b2 = ops.se_weights(op.module.fc1, op.module.fc2)
                hold = {"w1": w1, "b1": b1, "w2": w2, "b2": b2, "mod": op.module}
                self._se_holds = getattr(self, "_se_holds", []) + [hold]
                pooled = torch.empty((self.B, 32, op.src.C), dtype=torch.float32, device=dev)   # DYK_SE_MAX_SLABS partials
                gate = torch.empty((self.B, op.src.C), dtype=torch.float32, device=dev)
                op.gate = gate
                if op.gate_only:       # every consumer applies the gate while it loads (fuse_se_gates)
                    self.steps.append(_Call(ops.nhwc_se_gate, op.src.view, w1, b1, w2, b2, pooled, gate))
                    self.steps[-1].launches = 3
                else:
                    self.steps.append(_Call(ops.f32_se if self.f32 else ops.nhwc_se, op.src.view, op.out.view, w1, b1, w2, b2, pooled, gate))
            elif isinstance(op, YoloOp):
                m = op.module
                ny, nx = op.src.H, op.src.W
                if (m.nx, m.ny) != (nx, ny) or m.grid is None or m.anchor_vec.device != dev:
                    m.create_grids((nx, ny), dev)
                if not (op.src.f32 or self.f32):
                    raise nat.NativeError(f"layer {op.layer}: [yolo] must follow a linear conv without batch norm")
                p_out = torch.empty((self.B, m.na, ny, nx, m.no), dtype=torch.float32, device=dev)
                self.p_outs.append(p_out)
                anchor = m.anchor_vec.to(dev).float().contiguous()
                self.steps.append(_Call(ops.yolo_decode, op.src.view.buf, op.src.view.stride, p_out, self.io,
                                        N=self.B, ny=ny, nx=nx, na=m.na, no=m.no, anchor_vec=anchor, stride=m.stride,
                                        v4=m.bf_type == "yolov4", rows_total=rows_total, row_off=row_off, in_kind=2))
                if m.bf_type not in ("yolov3", "yolov4"):
                    raise TypeError("bounding box predication error")
                row_off += m.na * ny * nx
        row_off_ref[0] = row_off

    def _bind_add(self, op):
        m = op.module
        x = op.x
        n = len(op.others) + 1
        wall = None
        if m.weight:
            wall = torch.empty(n, dtype=torch.float32, device=self.device)
            self.steps.append(_Call(ops.fusion_weights, m.w, wall))
        cur = x.view
        for i, a in enumerate(op.others):
            last = i == len(op.others) - 1
            if a.C != x.C or (m.weight and n > 2):
                raise nat.NativeError(
                    f"layer {op.layer}: channel-mismatched / >2-way weighted shortcut is only available through "
                    "build_utils.layers.WeightedFeatureFusion.forward (no shipped cfg uses it)")
            dst = op.out.view if last else ops.new_view(self.B, x.H, x.W, x.C, self.dtype, self.device)
            if op.gate_of is not None and i == 0:
                self.steps.append(_Call(ops.nhwc_add, cur, a.view, dst, wall, gate=op.gate_of.gate))
            else:
                self.steps.append(_Call(ops.f32_add if self.f32 else ops.nhwc_add, cur, a.view, dst, wall))
            cur = dst

    def refresh_se(self):
        for h in getattr(self, "_se_holds", []):
            w1, b1, w2, b2 = ops.se_weights(h["mod"].fc1, h["mod"].fc2)
            h["w1"].copy_(w1); h["b1"].copy_(b1); h["w2"].copy_(w2); h["b2"].copy_(b2)

    # ---- execution ------------------------------------------------------------------------------
    def run_stems(self, x, y):
        stem = ops.f32_stem if self.f32 else ops.nhwc_stem
        for which, e, out, kw in self.stem_steps:
            src = x if which == 0 else y
            if (src.shape[2], src.shape[3]) != (self.H, self.W):     # model(x, y, input_size=...): resize fused into the stem
                if self.f32:
                    raise nat.NativeError("input_size: the fused bilinear resize exists in fp16 / bf16 only")
                stem(src, e["w"], e["scale"], e["bias"], out, resize_to=(self.H, self.W), **kw)
            else:
                stem(src, e["w"], e["scale"], e["bias"], out, **kw)

    def run_body(self):
        side = self.side_stream
        if side is None:
            for s in self.steps:
                if not isinstance(s, _Sync):
                    s()
            return
        main = torch.cuda.current_stream()
        side.wait_stream(main)                      # fork: the stems (caller's stream) precede both lanes
        streams = (main, side)
        for s in self.steps:
            if isinstance(s, _Sync):
                if s.kind == "record":
                    s.event.record(streams[s.lane])
                else:
                    streams[s.lane].wait_event(s.event)
            elif s.lane == 0:
                s()
            else:
                with torch.cuda.stream(side):
                    s()
        main.wait_stream(side)                      # join

    def capture(self):
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g):
                self.run_body()
            self.graph = g
        except Exception as exc:  # noqa: BLE001 - fall back to eager launches of the same native kernels
            import warnings
            warnings.warn(f"CUDA graph capture of the plan failed ({type(exc).__name__}: {exc}); "
                          "running the same native kernels as individual launches")
            self.graph_failed = True
            self.graph = None
            torch.cuda.synchronize()


class _Sync:
    """Cross-lane ordering marker in Plan.steps: record an event on a lane / make a lane wait for it."""
    launches = 0
    lane = 0

    def __init__(self, kind, event, lane):
        self.kind, self.event, self.lane = kind, event, lane

    def __call__(self):
        pass


class _Call:
    launches = 1
    lane = 0

    def __init__(self, fn, *args, **kw):
        self.fn, self.args, self.kw = fn, args, kw
        if fn is ops.nhwc_se:
            self.launches = 4

    def __call__(self):
        self.fn(*self.args, **self.kw)


class _ConvStep:
    launches = 1
    lane = 0

    def __init__(self, x, e, y, kw):
        self.x, self.e, self.y, self.kw = x, e, y, kw

    def __call__(self):
        e = self.e
        ops.nhwc_conv(self.x, e["w"], e["scale"], e["bias"], self.y, cout=e["conv"].out_channels, **self.kw)


class _GatedConvStep(_ConvStep):
    """1x1 convolution whose input carries a SqueezeExcitation gate: the gate is folded into per-image weights."""
    launches = 2

    def __init__(self, x, e, y, kw, se_op, wimg):
        super().__init__(x, e, y, kw)
        self.se_op, self.wimg = se_op, wimg

    def __call__(self):
        e = self.e
        ops.scale_weights_per_image(e["w"], self.se_op.gate, self.wimg)
        cout = e["conv"].out_channels
        ops.nhwc_conv(self.x, self.wimg, e["scale"], e["bias"], self.y, cout=cout,
                      w_image_stride=cout * e["conv"].in_channels, **self.kw)


class PlanCache:
    """Per-model cache of compiled plans, keyed on everything that changes the launch list."""

    def __init__(self, model):
        self.model = model
        self.plans = {}
        self.bank = WeightBank()
        self.last_plan = None
        self._tensors = None

    def invalidate(self):
        """Parameter storage moved (Module._apply: .to() / .half() / .cuda()): drop plans and packed weights."""
        self.plans.clear()
        self.bank = WeightBank()
        self._tensors = None

    def mark_stale(self):
        """Force a re-pack of every cached weight on the next run (same device buffers, so captured graphs stay valid)."""
        if self.bank.signature is not None:
            self.bank.signature = ("stale",)

    def run(self, x, y, input_size=None):
        model = self.model
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise ValueError("YOLO.forward expects (B, 3, H, W) tensors")
        ops._require_cuda(x, "YOLO.forward")
        if model.training:
            raise nat.NativeError("PlanCache.run is the eval-mode executor; training goes through dyk.train_plan")
        if y is not None and (y.shape != x.shape or y.device != x.device):
            raise ValueError("visible and LWIR batches must have the same shape and device")
        dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else model.compute_dtype
        if dtype not in (torch.float16, torch.bfloat16, torch.float32):
            raise nat.NativeError(f"compute dtype {dtype} is not supported (float16 / bfloat16 / float32)")

        def prep(t):
            if t is None:
                return None
            if t.dtype not in (torch.float32, torch.uint8):
                t = t.float()
            return t.contiguous()

        x, y = prep(x), prep(y)
        if y is not None and y.dtype != x.dtype:
            raise ValueError("visible and LWIR batches must have the same dtype")
        B, _, H, W = x.shape
        if input_size is not None:
            H, W = (int(input_size), int(input_size)) if isinstance(input_size, int) else (int(input_size[0]), int(input_size[1]))
        p0 = next(model.parameters())
        if p0.device != x.device:
            raise ValueError(f"model is on {p0.device} but the input is on {x.device}")
        if self._tensors is None:
            self._tensors = list(model.parameters()) + list(model.buffers())
        sig = WeightBank.model_signature(self._tensors)
        if self.bank.signature is not None and sig != self.bank.signature:
            # in-place updates (optimizer.step, load_state_dict): re-pack into the same device buffers
            self.bank.refresh()
            for pl in self.plans.values():
                pl.refresh_se()
        self.bank.signature = sig
        key = (B, H, W, dtype, y is not None, x.device)
        plan = self.plans.get(key)
        with torch.cuda.device(x.device):
            if plan is None:
                plan = Plan(model, self.bank, B, H, W, dtype, y is not None, x.device)
                self.plans[key] = plan
            self.last_plan = plan
            plan.run_stems(x, y)
            if model.use_cuda_graph and not plan.graph_failed and not torch.cuda.is_current_stream_capturing():
                if plan.graph is None:
                    plan.run_body()          # warm-up: one-time attribute / driver-entry-point setup
                    torch.cuda.synchronize()
                    plan.run_stems(x, y)
                    plan.capture()
                if plan.graph is not None:
                    plan.graph.replay()
                    nat.count_launches(plan.launches_per_forward - len(plan.stem_steps))
                else:
                    plan.run_body()
            else:
                plan.run_body()
        io = plan.io.clone()
        p = tuple(t.clone() for t in plan.p_outs)
        return io, p
