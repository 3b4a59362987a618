"""Generators for the Darknet ``.cfg`` model definitions the BASELINE configurations name.

The reference ships its model zoo as ~35 k lines of hand-written cfg text (config/*.cfg).  The cfg
format is the plugin format of the hot path, so this package has to be able to produce those model
definitions without carrying the reference's files: each function below *builds* the block list of
one architecture from its structure (Darknet-53 / CSPDarknet-53 / MobileNetV3 stages, fusion blocks,
SPP, FPN / PANet necks, heads) and ``materialize`` writes it as cfg text under a cache directory.
``tests/test_cfg_zoo.py`` checks block-for-block equality with the reference's files when the
reference tree is present and against a committed structural digest otherwise.

File names matter: the reference keys behaviour on substrings of the cfg *path* (models.py:124-126,131:
'yolov3'/'fpn'/'yolov-tiny' -> head strides [32,16,8]; 'yolov4' -> v4 box decode).
"""
from __future__ import annotations

import hashlib
import os
from pathlib import Path

ANCHORS_V3 = "16, 42, 22, 44, 20, 53, 29, 53, 26, 64, 29, 85, 34, 75, 41, 104, 59, 147"
ANCHORS_V3_VISIBLE = "16, 33, 18, 37, 19, 47, 23, 42, 20, 51, 28, 66, 37, 86, 45, 104, 58, 140"
ANCHORS_V4 = "16, 32, 18, 42, 22, 44, 22, 55, 30, 58, 27, 65, 34, 80, 43, 102, 62, 153"


class CfgBuilder:
    """Accumulates Darknet blocks; layer indices are those of the parsed list without ``[net]``."""

    def __init__(self, net: list):
        self.blocks = [("net", list(net))]

    @property
    def n(self) -> int:  # index the next layer will get
        return len(self.blocks) - 1

    def _add(self, kind, items):
        self.blocks.append((kind, items))
        return self.n - 1

    def conv(self, filters, size, stride=1, act="leaky", bn=True, groups=None):
        items = []
        if bn:
            items.append(("batch_normalize", 1))
        items += [("filters", filters), ("size", size), ("stride", stride), ("pad", 1)]
        if groups is not None:
            items.append(("groups", groups))
        items.append(("activation", act))
        return self._add("convolutional", items)

    def shortcut(self, frm, weighted=False):
        items = [("from", frm), ("activation", "linear")]
        if weighted:
            items.append(("weights_type", "1.0"))
        return self._add("shortcut", items)

    def route(self, *layers):
        return self._add("route", [("layers", ",".join(str(l) for l in layers))])

    def maxpool(self, size, stride=1):
        return self._add("maxpool", [("stride", stride), ("size", size)])

    def upsample(self, stride=2):
        return self._add("upsample", [("stride", stride)])

    def se(self, squeeze_factor=4):
        return self._add("se", [("squeeze_factor", squeeze_factor)])

    def dsconv(self, filters, stride=1, size=None):
        items = [("filters", filters)]
        if size is not None:
            items.append(("size", size))
        items.append(("stride", stride))
        return self._add("depthwiseconvolutional", items)

    def yolo(self, mask, anchors, extra=()):
        items = [("mask", ",".join(str(m) for m in mask)), ("anchors", anchors), ("classes", 1), ("num", 9),
                 ("jitter", ".3"), ("ignore_thresh", ".7"), ("truth_thresh", 1)]
        items += list(extra)
        return self._add("yolo", items)

    def text(self) -> str:
        out = []
        for kind, items in self.blocks:
            out.append("[%s]" % kind)
            out += ["%s = %s" % (k, v) for k, v in items]
            out.append("")
        return "\n".join(out)


# ----------------------------------------------------------------------------------------- [net]
def _net_v3(second_index=None):
    net = [("batch", 64), ("subdivisions", 16), ("width", 608), ("height", 608), ("channels", 3),
           ("momentum", "0.9"), ("decay", "0.0005"), ("angle", 0), ("saturation", "1.5"), ("exposure", "1.5"),
           ("hue", ".1")]
    if second_index is not None:
        net.append(("second_index", second_index))
    net += [("learning_rate", "0.001"), ("burn_in", 1000), ("max_batches", 500200), ("policy", "steps"),
            ("steps", "400000,450000"), ("scales", ".1,.1")]
    return net


def _net_v4(second_index, commented_size=True):
    size_notes = [(";width", 608), (";height", 608)] if commented_size else []
    return [("batch", 64), ("subdivisions", 8), ("width", 512), ("height", 512)] + size_notes + [
            ("channels", 3), ("momentum", "0.949"), ("decay", "0.0005"), ("angle", 0),
            ("saturation", "1.5"), ("exposure", "1.5"), ("hue", ".1"), ("second_index", second_index),
            ("learning_rate", "0.0013"), ("burn_in", 1000), ("max_batches", 500500), ("policy", "steps"),
            ("steps", "400000,450000"), ("scales", ".1,.1"), ("mosaic", 1)]


# ----------------------------------------------------------------------------------------- Darknet-53
def _darknet53(b: CfgBuilder):
    """Stem + 5 stages of (1x1 -> 3x3 -> add) residual blocks; returns the last layer of stages 3,4,5."""
    b.conv(32, 3)
    taps = []
    for filters, blocks in ((64, 1), (128, 2), (256, 8), (512, 8), (1024, 4)):
        b.conv(filters, 3, stride=2)
        for _ in range(blocks):
            b.conv(filters // 2, 1)
            b.conv(filters, 3)
            last = b.shortcut(-3)
        taps.append(last)
    return taps[2], taps[3], taps[4]


def _spp(b: CfgBuilder):
    b.maxpool(5)
    b.route(-2)
    b.maxpool(9)
    b.route(-4)
    b.maxpool(13)
    b.route(-1, -3, -5, -6)


def _yolov3_spp_neck(b: CfgBuilder, up1_from: int, up2_from: int, anchors: str = ANCHORS_V3):
    """YOLOv3-SPP neck and the three heads (stride 32, 16, 8); up*_from are relative route offsets."""
    v3_yolo_extra = [("random", 1)]
    b.conv(512, 1); b.conv(1024, 3); b.conv(512, 1)
    _spp(b)
    b.conv(512, 1); b.conv(1024, 3); b.conv(512, 1); b.conv(1024, 3)
    b.conv(18, 1, act="linear", bn=False)
    b.yolo((6, 7, 8), anchors, v3_yolo_extra)
    for width, mask, up_from in ((256, (3, 4, 5), up1_from), (128, (0, 1, 2), up2_from)):
        b.route(-4)
        b.conv(width, 1)
        b.upsample(2)
        b.route(-1, up_from)
        for _ in range(3):
            b.conv(width, 1)
            b.conv(width * 2, 3)
        b.conv(18, 1, act="linear", bn=False)
        b.yolo(mask, anchors, v3_yolo_extra)


def kaist_yolov3() -> str:
    """Visible-only YOLOv3-SPP (reference config/kaist_yolov3.cfg)."""
    b = CfgBuilder(_net_v3())
    _darknet53(b)
    _yolov3_spp_neck(b, up1_from=61, up2_from=36, anchors=ANCHORS_V3_VISIBLE)
    return b.text()


def kaist_dyolov3_add_sl() -> str:
    """Dual Darknet-53 + self-learned weighted-add fusion at strides 8/16/32
    (reference config/kaist_dyolov3_add_sl.cfg)."""
    b = CfgBuilder(_net_v3(second_index=75))
    v8, v16, v32 = _darknet53(b)
    assert b.n == 75
    l8, l16, l32 = _darknet53(b)
    # fusion: route(visible tap) -> weighted shortcut(lwir tap) -> 3x3 conv
    b.route(v8 - b.n); b.shortcut(l8, weighted=True); b.conv(256, 3)
    b.route(v16); b.shortcut(l16, weighted=True); b.conv(512, 3)
    b.route(v32); b.shortcut(l32, weighted=True); b.conv(1024, 3)
    _yolov3_spp_neck(b, up1_from=-22, up2_from=-37)
    return b.text()


# ----------------------------------------------------------------------------------------- CSPDarknet-53
def _csp_stage(b: CfgBuilder, filters: int, blocks: int, first: bool = False):
    """Downsample + cross-stage-partial stage with `blocks` residual units (Mish)."""
    m = "mish"
    b.conv(filters, 3, stride=2, act=m)
    half = filters if first else filters // 2
    b.conv(half, 1, act=m)
    b.route(-2)
    b.conv(half, 1, act=m)
    for _ in range(blocks):
        b.conv(filters // 2 if first else half, 1, act=m)
        b.conv(half, 3, act=m)
        b.shortcut(-3)
    b.conv(half, 1, act=m)
    b.route(-1, -(3 * blocks + 4))
    return b.conv(filters, 1, act=m)


def _csp_front(b: CfgBuilder):
    """Stem + CSP stages 1..3 (to stride 8, 256 channels)."""
    b.conv(32, 3, act="mish")
    _csp_stage(b, 64, 1, first=True)
    _csp_stage(b, 128, 2)
    return _csp_stage(b, 256, 8)


def _fsnet_fuse(b: CfgBuilder, a: int, c: int, filters: int):
    """route(a, c) -> 3x3 conv (Mish) -> SE: the concat-SE fusion block; returns the SE layer index."""
    b.route(a, c)
    b.conv(filters, 3, act="mish")
    return b.se(4)


def kaist_dyolov4_fshare_global_concat_se3() -> str:
    """Dual CSPDarknet-53 with FSNet (shared fused features re-injected into both branches), concat-SE
    fusion at strides 8/16/32, SPP + PANet (reference config/kaist_dyolov4_fshare_global_concat_se3.cfg)."""
    b = CfgBuilder(_net_v4(second_index=55))
    v3 = _csp_front(b)
    assert b.n == 55
    l3 = _csp_front(b)
    f1 = _fsnet_fuse(b, v3, l3, 256)
    b.shortcut(v3, weighted=True)
    v4 = _csp_stage(b, 512, 8)
    b.route(f1)
    b.shortcut(l3, weighted=True)
    l4 = _csp_stage(b, 512, 8)
    f2 = _fsnet_fuse(b, v4, l4, 512)
    b.shortcut(v4, weighted=True)
    v5 = _csp_stage(b, 1024, 4)
    b.route(f2)
    b.shortcut(l4, weighted=True)
    l5 = _csp_stage(b, 1024, 4)
    _fsnet_fuse(b, v5, l5, 1024)
    # SPP neck
    b.conv(512, 1); b.conv(1024, 3); b.conv(512, 1)
    _spp(b)
    b.conv(512, 1); b.conv(1024, 3); b.conv(512, 1)
    # PANet top-down
    for width, lateral in ((256, f2), (128, f1)):
        b.conv(width, 1)
        b.upsample(2)
        b.route(lateral)
        b.conv(width, 1)
        b.route(-1, -3)
        b.conv(width, 1); b.conv(width * 2, 3); b.conv(width, 1); b.conv(width * 2, 3); b.conv(width, 1)
    v4_extra = lambda sxy, rnd=False: ([("random", 1)] if rnd else []) + [("scale_x_y", sxy), ("iou_thresh", "0.213"), ("cls_normalizer", "1.0"),
                            ("iou_normalizer", "0.07"), ("iou_loss", "ciou"), ("nms_kind", "greedynms"),
                            ("beta_nms", "0.6")]
    b.conv(256, 3)
    b.conv(18, 1, act="linear", bn=False)
    b.yolo((0, 1, 2), ANCHORS_V4, v4_extra("1.2"))
    # PANet bottom-up
    for width, mask, back, sxy in ((256, (3, 4, 5), -16, "1.1"), (512, (6, 7, 8), -37, "1.05")):
        last = mask[0] == 6
        b.route(-4)
        b.conv(width, 3, stride=2)
        b.route(-1, back)
        b.conv(width, 1); b.conv(width * 2, 3); b.conv(width, 1); b.conv(width * 2, 3); b.conv(width, 1)
        b.conv(width * 2, 3)
        b.conv(18, 1, act="linear", bn=False)
        b.yolo(mask, ANCHORS_V4, v4_extra(sxy, rnd=last))
    return b.text()


# ----------------------------------------------------------------------------------------- MobileNetV3
def _bneck(b: CfgBuilder, exp: int, k: int, out: int, se: bool, act: str, stride: int = 1, res: bool = False):
    """MobileNetV3 bottleneck as the reference writes it: 1x1 expand (carries the stride) -> depthwise kxk
    -> [SE] -> 1x1 linear projection -> [residual]."""
    b.conv(exp, 1, stride=stride, act=act)
    b.conv(exp, k, groups=exp, act=act)
    if se:
        b.se(4)
    last = b.conv(out, 1, act="linear")
    if res:
        last = b.shortcut(-5 if se else -4)
    return last


def _mnv3_front(b: CfgBuilder):
    b.conv(16, 3, stride=2, act="hard-swish")
    b.conv(16, 3, groups=16, act="relu")
    b.conv(16, 1, act="linear")
    _bneck(b, 64, 3, 24, False, "relu", stride=2)
    _bneck(b, 72, 3, 24, False, "relu", res=True)
    _bneck(b, 72, 5, 40, True, "relu", stride=2)
    _bneck(b, 120, 5, 40, True, "relu", res=True)
    return _bneck(b, 120, 5, 40, True, "relu", res=True)


def _mnv3_stage4(b: CfgBuilder):
    hs = "hard-swish"
    _bneck(b, 240, 3, 80, False, hs, stride=2)
    _bneck(b, 200, 3, 80, False, hs, res=True)
    _bneck(b, 184, 3, 80, False, hs, res=True)
    _bneck(b, 184, 3, 80, False, hs, res=True)
    _bneck(b, 480, 3, 112, True, hs)
    return _bneck(b, 672, 3, 112, True, hs, res=True)


def _mnv3_stage5(b: CfgBuilder):
    hs = "hard-swish"
    _bneck(b, 672, 5, 160, True, hs, stride=2)
    _bneck(b, 960, 5, 160, True, hs, res=True)
    return _bneck(b, 960, 5, 160, True, hs, res=True)


def _cse_light(b: CfgBuilder, a: int, c: int, ch: int):
    """Lightweight concat-SE fusion: route(a, c) -> depthwise 3x3 (ReLU6) -> SE -> 1x1 linear.
    Returns (depthwise layer, projection layer)."""
    b.route(a, c)
    dw = b.conv(2 * ch, 3, groups=2 * ch, act="relu6")
    b.se(4)
    return dw, b.conv(ch, 1, act="linear")


def kaist_dyolov4_mobilenetv3_fshare_global_cse3() -> str:
    """Dual MobileNetV3-large backbones with FSNet re-injection and light concat-SE fusion, SPP + PANet neck
    built from depthwise-separable convs (reference config/kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg)."""
    b = CfgBuilder(_net_v4(second_index=24, commented_size=False))
    v3 = _mnv3_front(b)
    assert b.n == 24
    l3 = _mnv3_front(b)
    lwir_se3 = l3 - 2                      # SE output inside the last LWIR bottleneck (PANet lateral tap)
    _, f1 = _cse_light(b, v3, l3, 40)
    b.shortcut(v3, weighted=True)
    v4 = _mnv3_stage4(b)
    b.route(f1)
    b.shortcut(l3, weighted=True)
    l4 = _mnv3_stage4(b)
    dw2, f2 = _cse_light(b, v4, l4, 112)
    b.shortcut(v4, weighted=True)
    v5 = _mnv3_stage5(b)
    b.route(f2)
    b.shortcut(l4, weighted=True)
    l5 = _mnv3_stage5(b)
    _cse_light(b, v5, l5, 160)
    r6 = "relu6"
    b.conv(512, 1, act=r6); b.dsconv(1024); b.conv(512, 1, act=r6)
    _spp(b)
    b.conv(512, 1, act=r6); b.dsconv(1024); b.conv(512, 1, act=r6)
    for width, lateral, n_pairs in ((256, dw2, 2), (128, lwir_se3, 2)):
        b.conv(width, 1, act=r6)
        b.upsample(2)
        b.route(lateral)
        b.conv(width, 1, act=r6)
        b.route(-1, -3)
        for _ in range(n_pairs):
            b.conv(width, 1, act=r6)
            b.dsconv(width * 2)
        b.conv(width, 1, act=r6)
    v4_extra = lambda sxy, rnd=False: ([("random", 1)] if rnd else []) + [
        ("scale_x_y", sxy), ("iou_thresh", "0.213"), ("cls_normalizer", "1.0"), ("iou_normalizer", "0.07"),
        ("iou_loss", "ciou"), ("nms_kind", "greedynms"), ("beta_nms", "0.6")]
    b.dsconv(256)
    b.conv(18, 1, act="linear", bn=False)
    b.yolo((0, 1, 2), ANCHORS_V4, v4_extra("1.2"))
    for width, mask, back, sxy, first_act in ((256, (3, 4, 5), -16, "1.1", r6), (512, (6, 7, 8), -37, "1.05", "leaky")):
        b.route(-4)
        b.dsconv(width, stride=2)
        b.route(-1, back)
        b.conv(width, 1, act=first_act); b.dsconv(width * 2)
        b.conv(width, 1, act=r6); b.dsconv(width * 2)
        b.conv(width, 1, act=r6); b.dsconv(width * 2)
        b.conv(18, 1, act="linear", bn=False)
        b.yolo(mask, ANCHORS_V4, v4_extra(sxy, rnd=mask[0] == 6))
    return b.text()


ZOO = {
    "kaist_yolov3.cfg": kaist_yolov3,
    "kaist_dyolov3_add_sl.cfg": kaist_dyolov3_add_sl,
    "kaist_dyolov4_fshare_global_concat_se3.cfg": kaist_dyolov4_fshare_global_concat_se3,
    "kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg": kaist_dyolov4_mobilenetv3_fshare_global_cse3,
}


def cache_dir() -> Path:
    d = os.environ.get("DYK_CFG_DIR")
    p = Path(d) if d else Path(__file__).resolve().parent.parent / "_generated_cfg"
    p.mkdir(parents=True, exist_ok=True)
    return p


def materialize(name: str) -> str:
    """Writes the generated cfg (idempotent) and returns its path; `name` keeps the reference's file name."""
    if name not in ZOO:
        raise KeyError(f"no generator for {name}; known: {sorted(ZOO)}")
    text = ZOO[name]()
    path = cache_dir() / name
    if not path.exists() or path.read_text(encoding="utf-8") != text:
        tmp = path.with_suffix(".tmp%d" % os.getpid())
        tmp.write_text(text, encoding="utf-8")
        os.replace(tmp, path)
    return str(path)


def structural_digest(blocks: list) -> str:
    """sha256 over the parsed block list (net dict included), independent of cfg text formatting."""
    h = hashlib.sha256()
    for blk in blocks:
        for k in sorted(blk):
            v = blk[k]
            v = v.tolist() if hasattr(v, "tolist") else v
            h.update(repr((k, v)).encode())
        h.update(b"|")
    return h.hexdigest()
