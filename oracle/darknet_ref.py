"""ORACLE — TEST INFRASTRUCTURE ONLY.  Plain PyTorch fp32 restatement of the reference's hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; nothing under double-yolo-kaist_b200/ does (tests/test_boundary.py greps for it).

What it restates (paths relative to the reference repo, Ye-zixiao/Double-YOLO-Kaist @ b4f50aa):
  * build_utils/parse_config.py:5-65   parse_model_cfg           -> parse_cfg
  * models.py:7-155                    create_modules            -> DarknetRef.__init__ (parameter table only)
  * models.py:279-315                  YOLO.forward              -> DarknetRef.forward
  * models.py:218-258                  YOLOLayer.forward         -> yolo_layer
  * build_utils/layers.py:32-44        FeatureConcat             -> inline in forward ("route")
  * build_utils/layers.py:47-85        WeightedFeatureFusion     -> weighted_fusion
  * build_utils/layers.py:148-172      Inception                 -> inception
  * build_utils/layers.py:175-190      SqueezeExcitation         -> squeeze_excitation
  * build_utils/layers.py:218-234      DepthwiseSeparableConv2d  -> inline in forward
  * build_utils/utils.py:50-57,387-464 xywh2xyxy, non_max_suppression (+ torchvision.ops.nms) -> nms_ref.py

The arithmetic of the reference lives in PyTorch library calls (nn.Conv2d, nn.BatchNorm2d, activations,
nn.MaxPool2d, nn.Upsample, torch.cat ...), so the restatement calls the same torch.nn.functional ops in
fp32; it differs from the reference in structure (a functional interpreter over a flat parameter dict
keyed by the reference's state_dict names), not in numerics.

PINNING: the reference has no tests or golden vectors.  This oracle is pinned against
outputs of the reference itself, generated in the build container by tests/golden/make_golden.py (which
imports /root/reference) and committed under tests/golden/*.npz; tests/test_oracle_golden.py replays them.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

_SUPPORTED = {
    'type', 'batch_normalize', 'filters', 'size', 'stride', 'pad', 'activation', 'layers', 'groups', 'from', 'mask',
    'anchors', 'classes', 'num', 'jitter', 'ignore_thresh', 'truth_thresh', 'random', 'stride_x', 'stride_y',
    'weights_type', 'weights_normalization', 'scale_x_y', 'beta_nms', 'nms_kind', 'iou_loss', 'iou_normalizer',
    'cls_normalizer', 'iou_thresh', 'probability', 'max_delta', 'atoms', 'na', 'nc', 'squeeze_factor', 'n1x1',
    'n3x3_reduce', 'n3x3', 'n5x5_reduce', 'n5x5', 'pool_proj'}


def parse_cfg(path: str) -> list:
    """parse_config.py:5-65 — list of dicts, first one is [net]."""
    if not path.endswith(".cfg") or not os.path.exists(path):
        raise FileNotFoundError("the cfg file not exist...")
    with open(path, "r", encoding="utf-8") as f:
        lines = f.read().split("\n")
    lines = [ln.strip() for ln in lines if ln and not ln.startswith("#")]
    out = []
    for ln in lines:
        if ln.startswith("["):
            out.append({"type": ln[1:-1].strip()})
            if out[-1]["type"] == "convolutional":
                out[-1]["batch_normalize"] = 0
            continue
        key, val = (s.strip() for s in ln.split("="))
        if key == "anchors":
            out[-1][key] = np.array([float(x) for x in val.replace(" ", "").split(",")]).reshape((-1, 2))
        elif key in ("from", "layers", "mask") or (key == "size" and "," in val):
            out[-1][key] = [int(x) for x in val.split(",")]
        elif val.isnumeric():
            out[-1][key] = int(val)
        else:
            out[-1][key] = val
    for blk in out[1:]:
        bad = [k for k in blk if k not in _SUPPORTED]
        if bad:
            raise ValueError("Unsupported fields:{} in cfg".format(bad[0]))
    return out


def activation(x, name):
    """models.py:51-64 — the nn modules the reference instantiates, as functionals."""
    if name == "leaky":
        return F.leaky_relu(x, 0.1)
    if name == "mish":
        return F.mish(x)
    if name == "relu":
        return F.relu(x)
    if name == "relu6":
        return F.relu6(x)
    if name == "hard-swish":
        return F.hardswish(x)
    if name == "hard-sigmoid":
        return F.hardsigmoid(x)
    return x


def make_divisible(v, d):
    return math.ceil(v / d) * d


class DarknetRef:
    """Functional interpreter of a Darknet cfg over a flat ``state`` dict (reference state_dict names)."""

    def __init__(self, cfg_path: str):
        self.cfg = cfg_path
        blocks = parse_cfg(cfg_path)
        self.net = blocks[0]
        self._round = None
        self.defs = blocks[1:]
        self.shapes = OrderedDict()   # state_dict name -> shape
        self.meta = []                # per layer: dict with derived constants
        self._plan_params()

    # -- models.py:7-155 ---------------------------------------------------------------------------
    def _plan_params(self):
        out_filters = [3]
        second = self.net.get("second_index")
        yolo_index = -1
        sh = self.shapes

        def conv_bn(prefix_conv, prefix_bn, cin, cout, k, groups, bn, bias):
            sh[prefix_conv + ".weight"] = (cout, cin // groups, k, k)
            if bias:
                sh[prefix_conv + ".bias"] = (cout,)
            if bn:
                for nm in ("weight", "bias", "running_mean", "running_var"):
                    sh[prefix_bn + "." + nm] = (cout,)
                sh[prefix_bn + ".num_batches_tracked"] = ()

        def cba(prefix, cin, cout, k):
            conv_bn(prefix + ".conv.0", prefix + ".conv.1", cin, cout, k, 1, True, False)

        for i, d in enumerate(self.defs):
            t = d["type"]
            m = {}
            filters = out_filters[-1]
            pre = f"module_list.{i}"
            if t == "convolutional":
                bn = d["batch_normalize"]
                filters = d["filters"]
                cin = 3 if (second is not None and i == second) else out_filters[-1]
                g = d.get("groups", 1)
                conv_bn(pre + ".Conv2d", pre + ".BatchNorm2d", cin, filters, d["size"], g, bool(bn), not bn)
                m.update(k=d["size"], stride=d["stride"], pad=d["size"] // 2 if d["pad"] else 0, groups=g, bn=bool(bn))
            elif t == "depthwiseconvolutional":
                cin = out_filters[-1]
                filters = d["filters"]
                ks = d.get("size", 3)
                conv_bn(pre + ".conv.0", pre + ".conv.1", cin, cin, ks, cin, True, False)
                conv_bn(pre + ".conv.3", pre + ".conv.4", cin, filters, 1, 1, True, False)
                m.update(k=ks, stride=d["stride"])
            elif t == "inception":
                cin = out_filters[-1]
                cba(pre + ".branch1.0", cin, d["n1x1"], 1)
                cba(pre + ".branch2.0", cin, d["n3x3_reduce"], 1)
                cba(pre + ".branch2.1", d["n3x3_reduce"], d["n3x3"], 3)
                cba(pre + ".branch3.0", cin, d["n5x5_reduce"], 1)
                cba(pre + ".branch3.1", d["n5x5_reduce"], d["n5x5"], 3)
                cba(pre + ".branch3.2", d["n5x5"], d["n5x5"], 3)
                cba(pre + ".branch4.1", cin, d["pool_proj"], 1)
            elif t == "se":
                c = out_filters[-1]
                sq = make_divisible(c // d["squeeze_factor"], 8)
                sh[pre + ".fc1.weight"] = (sq, c, 1, 1)
                sh[pre + ".fc1.bias"] = (sq,)
                sh[pre + ".fc2.weight"] = (c, sq, 1, 1)
                sh[pre + ".fc2.bias"] = (c,)
            elif t == "route":
                layers = d["layers"]
                filters = sum(out_filters[l + 1 if l > 0 else l] for l in layers)
                m["layers"] = [i + l if l < 0 else l for l in layers]
            elif t == "shortcut":
                m["layers"] = [i + l if l < 0 else l for l in d["from"]]
                m["weight"] = "weights_type" in d
                if m["weight"]:
                    sh[pre + ".w"] = (len(m["layers"]) + 1,)
            elif t == "yolo":
                yolo_index += 1
                stride = [8, 16, 32, 64, 128]
                if any(x in self.cfg for x in ["yolov-tiny", "fpn", "yolov3"]):
                    stride = [32, 16, 8]
                m["stride"] = stride[yolo_index]
                m["anchors"] = torch.tensor(d["anchors"][d["mask"]], dtype=torch.float32)
                m["nc"] = d["classes"]
                m["v4"] = "yolov4" in self.cfg
            self.meta.append(m)
            out_filters.append(filters)

    # -- building blocks ---------------------------------------------------------------------------
    @staticmethod
    def _bn(x, st, prefix, training, stats_momentum):
        w, b = st[prefix + ".weight"], st[prefix + ".bias"]
        rm, rv = st[prefix + ".running_mean"], st[prefix + ".running_var"]
        if training:
            nbt = st.get(prefix + ".num_batches_tracked")
            if nbt is not None:
                nbt += 1
            mom = stats_momentum
            if mom is None:  # nn.BatchNorm2d(momentum=None): cumulative moving average
                mom = 1.0 / float(nbt.item())
            return F.batch_norm(x, rm, rv, w, b, True, mom, 1e-5)
        return F.batch_norm(x, rm, rv, w, b, False, 0.0, 1e-5)

    def _conv_bn_act(self, x, st, pconv, pbn, stride, pad, groups, act, training, mom, gate=None):
        w = st[pconv + ".weight"]
        if self._round is not None and groups == 1 and w.shape[1] > 4:
            w = w.to(self._round).float()   # dense tensor-core convs hold their weights in the 16-bit dtype
        if gate is not None:
            # storage-rounding model of a SqueezeExcitation gate folded into the consuming 1x1 convolution (dyk/plan.py
            # fuse_se_gates): image n is convolved with its own weights round(W * gate[n]); x is the [se] block's INPUT
            outs = []
            for n in range(x.shape[0]):
                wn = w * gate[n].reshape(1, -1, 1, 1)
                if self._round is not None:
                    wn = wn.to(self._round).float()
                outs.append(F.conv2d(x[n:n + 1], wn, st.get(pconv + ".bias"), stride, pad, 1, groups))
            x = torch.cat(outs, 0)
            if pbn is not None:
                x = self._bn(x, st, pbn, training, mom)
            return activation(x, act)
        elif (self._round is not None and groups == 1 and tuple(w.shape) == (32, 3, 3, 3) and stride == 1 and pad == 1):
            # the 3 -> 32 stem runs on the tensor cores too (csrc/conv_stem_tc.cu): frames and weights are rounded to the
            # 16-bit type when the im2col rows are built; other stem shapes stay on the fp32 CUDA-core kernel
            w = w.to(self._round).float()
            x = x.to(self._round).float()
        x = F.conv2d(x, w, st.get(pconv + ".bias"), stride, pad, 1, groups)
        if pbn is not None and training and self._round is not None:
            x = x.to(self._round).float()   # the training plan stores the pre-BN conv output in 16 bits
        if pbn is not None:
            x = self._bn(x, st, pbn, training, mom)
        return activation(x, act)

    def _cba(self, x, st, prefix, k, training, mom):  # layers.py:88-122 with defaults stride 1, leaky, bn
        return self._conv_bn_act(x, st, prefix + ".conv.0", prefix + ".conv.1", 1, k // 2 if k == 3 else 0, 1,
                                 "leaky", training, mom)

    # -- models.py:279-315 -------------------------------------------------------------------------
    def forward(self, st: dict, x: torch.Tensor, y: torch.Tensor = None, training: bool = False,
                bn_momentum=0.1, keep_layers: bool = False, round_dtype=None, unrounded=(), teacher=None, gated=None,
                teacher_gates=None):
        """Returns what YOLO.forward returns: train -> [p...]; eval -> (cat(io, 1), (p...)).
        With keep_layers=True also returns the list of every layer's output (for per-layer checks).

        round_dtype (torch.float16 / torch.bfloat16) turns this into the *storage-rounding model* of the native
        path: arithmetic stays fp32, but dense-conv weights and every layer output are rounded to that dtype,
        except the layers listed in `unrounded` (outputs the native plan never materialises because they are
        fused into their consumer, and the fp32 head logits).

        teacher ({layer index: NCHW fp32 tensor}) turns the run into a *teacher-forced, layer-by-layer* check: after
        layer i has been computed (and recorded in the keep_layers list) its output is replaced by teacher[i] — the
        tensor the native plan actually produced — so every layer is evaluated on exactly the inputs the native
        kernel saw.  Rounding is a non-linear amplifier (a 1e-7 accumulation-order difference flips a 16-bit
        rounding somewhere, and within ~5 layers two correct implementations sit a full rounding-noise floor
        apart), so end-to-end comparisons can only bound drift; the teacher-forced comparison is tight (<= 1-2 ulp
        of the storage type per layer) for every layer of the real network at its real shape."""
        # gated ({conv layer: se layer}, from the native plan): those 1x1 convolutions read the [se] block's input and fold its
        # gate into per-image weights instead of reading the gated tensor (same mathematics, rounding at another place)
        gated = gated or {}
        se_in, se_gate = {}, {}
        self._round = round_dtype
        rnd = (lambda t: t.to(round_dtype).float()) if round_dtype is not None else (lambda t: t)
        di = "second_index" in self.net and y is not None
        yolo_out, out, every = [], [], []
        routed = set()
        for i, (d, m) in enumerate(zip(self.defs, self.meta)):
            if d["type"] in ("route", "shortcut"):
                routed.update(m["layers"])
            if d["type"] == "convolutional" and not m["bn"]:
                routed.add(i)
        for i, (d, m) in enumerate(zip(self.defs, self.meta)):
            t = d["type"]
            pre = f"module_list.{i}"
            if t == "convolutional":
                src = y if (di and i == self.net["second_index"]) else x
                g = None
                if i in gated:
                    src, g = se_in[gated[i]], se_gate[gated[i]]
                x = self._conv_bn_act(src, st, pre + ".Conv2d", pre + ".BatchNorm2d" if m["bn"] else None, m["stride"],
                                      m["pad"], m["groups"], d["activation"], training, bn_momentum, gate=g)
            elif t == "depthwiseconvolutional":
                c = x.shape[1]
                x = rnd(self._conv_bn_act(x, st, pre + ".conv.0", pre + ".conv.1", m["stride"], 1, c, "relu6", training,
                                          bn_momentum))
                x = self._conv_bn_act(x, st, pre + ".conv.3", pre + ".conv.4", 1, 0, 1, "relu6", training, bn_momentum)
            elif t == "inception":
                b1 = self._cba(x, st, pre + ".branch1.0", 1, training, bn_momentum)
                b2 = self._cba(rnd(self._cba(x, st, pre + ".branch2.0", 1, training, bn_momentum)), st, pre + ".branch2.1", 3,
                               training, bn_momentum)
                b3 = rnd(self._cba(x, st, pre + ".branch3.0", 1, training, bn_momentum))
                b3 = rnd(self._cba(b3, st, pre + ".branch3.1", 3, training, bn_momentum))
                b3 = self._cba(b3, st, pre + ".branch3.2", 3, training, bn_momentum)
                b4 = self._cba(F.max_pool2d(x, 3, 1, 1), st, pre + ".branch4.1", 1, training, bn_momentum)
                x = torch.cat([b1, b2, b3, b4], 1)
            elif t == "se":  # layers.py:184-190
                s = F.adaptive_avg_pool2d(x, (1, 1))
                s = F.relu(F.conv2d(s, st[pre + ".fc1.weight"], st[pre + ".fc1.bias"]))
                s = F.hardsigmoid(F.conv2d(s, st[pre + ".fc2.weight"], st[pre + ".fc2.bias"]))
                if teacher_gates is not None and i in teacher_gates:
                    # teacher forcing of a gate-only block: its consumers are checked against the gate the native run used
                    # (the gate itself is checked here, to fp32 accumulation order)
                    tg = teacher_gates[i].reshape(s.shape)
                    assert float((tg - s).abs().max()) < 2e-5, ("SE gate", i, float((tg - s).abs().max()))
                    s = tg
                se_in[i], se_gate[i] = x, s.flatten(1)
                x = s * x
            elif t == "maxpool":
                k = d["size"]
                x = F.max_pool2d(x, k, d["stride"], (k - 1) // 2)
            elif t == "upsample":
                x = F.interpolate(x, scale_factor=d["stride"], mode="nearest")
            elif t == "dropout":
                x = F.dropout(x, float(d["probability"]), training)
            elif t == "route":  # layers.py:44
                ls = m["layers"]
                x = torch.cat([out[l] for l in ls], 1) if len(ls) > 1 else out[ls[0]]
            elif t == "shortcut":  # layers.py:63-85
                x = weighted_fusion(x, [out[l] for l in m["layers"]], st[pre + ".w"] if m["weight"] else None)
            elif t == "yolo":
                yolo_out.append(yolo_layer(x, m["anchors"], m["stride"], m["nc"], m["v4"], training))
            if round_dtype is not None and i not in unrounded and t != "yolo":
                x = rnd(x)
            if keep_layers:
                every.append(x)
            if teacher is not None and i in teacher and t != "yolo":
                x = teacher[i]
            out.append(x if i in routed else None)
        if training:
            res = yolo_out
        else:
            io, p = zip(*yolo_out)
            res = (torch.cat(io, 1), p)
        return (res, every) if keep_layers else res


def weighted_fusion(x, others, w_param):
    """layers.py:63-85."""
    n = len(others) + 1
    if w_param is not None:
        w = torch.sigmoid(w_param) * (2 / n)
        x = x * w[0]
    nx = x.shape[1]
    for i, a in enumerate(others):
        if w_param is not None:
            a = a * w[i + 1]
        na = a.shape[1]
        if nx == na:
            x = x + a
        elif nx > na:
            x = torch.cat([x[:, :na] + a, x[:, na:]], 1)   # the reference writes the slice in place
        else:
            x = x + a[:, :nx]
    return x


def yolo_layer(p, anchors, stride, nc, v4, training):
    """models.py:218-258."""
    bs, _, ny, nx = p.shape
    na, no = anchors.shape[0], nc + 5
    p = p.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
    if training:
        return p
    anchor_wh = (anchors.to(p.device) / stride).view(1, na, 1, 1, 2)
    yv, xv = torch.meshgrid([torch.arange(ny, device=p.device), torch.arange(nx, device=p.device)], indexing="ij")
    grid = torch.stack((xv, yv), 2).view(1, 1, ny, nx, 2).float()
    if not v4:
        io = p.clone()
        io[..., :2] = torch.sigmoid(io[..., :2]) + grid
        io[..., 2:4] = torch.exp(io[..., 2:4]) * anchor_wh
        io[..., :4] *= stride
        torch.sigmoid_(io[..., 4:])
    else:
        io = p.sigmoid()
        io[..., :2] = (io[..., :2] * 2. - 0.5 + grid)
        io[..., 2:4] = (io[..., 2:4] * 2) ** 2 * anchor_wh
        io[..., :4] *= stride
    return io.view(bs, -1, no), p
