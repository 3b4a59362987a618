"""Per-kernel parity on a B200: every native kernel against the same op in plain PyTorch fp32 (the ops the
reference dispatches to).  Inputs and weights are first rounded to the kernel's storage dtype so the only
differences left are accumulation order and the final fp16/bf16 rounding; tolerances are written per test."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _setup(native_lib):
    assert torch.cuda.is_available(), "GPU tests need a B200"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from dyk import _native
    _native.check_device()


def _q(t, dtype):
    return t.to(dtype).float()


def _close(got, want, rtol, atol, what=""):
    got, want = got.float().cpu(), want.float().cpu()
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = err > tol
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())}/{bad.numel()} elements off, max err "
                                 f"{float(err.max()):.4g} at ref {float(want.flatten()[err.argmax()]):.4g}")


def _act_ref(x, act):
    return {"linear": lambda v: v, "leaky": lambda v: F.leaky_relu(v, 0.1), "mish": F.mish, "relu": F.relu,
            "relu6": F.relu6, "hard-swish": F.hardswish, "hard-sigmoid": F.hardsigmoid}[act](x)


CONV_CASES = [
    # (N, Cin, H, W, Cout, k, stride, act, residual, upsample, bn)
    (2, 64, 16, 20, 64, 1, 1, "leaky", False, False, True),
    (2, 64, 16, 20, 128, 3, 1, "leaky", False, False, True),
    (2, 128, 16, 20, 128, 3, 1, "mish", True, False, True),
    (1, 32, 32, 40, 64, 3, 2, "leaky", False, False, True),      # Cin 32 stride 2: pixel-pair K blocks
    (3, 32, 34, 46, 64, 3, 2, "mish", False, False, True),      # the same with ragged tiles
    (2, 32, 31, 42, 128, 3, 2, "leaky", False, False, True),    # odd height
    (2, 32, 30, 41, 64, 3, 2, "leaky", False, False, True),     # odd width: falls back to the parity-plane path
    (2, 64, 17, 23, 96, 3, 2, "mish", False, False, True),      # odd sizes: parity planes + edge clipping
    (2, 256, 8, 10, 512, 3, 1, "leaky", True, False, True),
    (2, 512, 4, 5, 256, 1, 1, "leaky", False, True, True),      # fused 2x upsample
    (1, 1024, 4, 5, 18, 1, 1, "linear", False, False, False),   # detection head (fp32 out path is in the model test)
    (3, 24, 12, 12, 72, 1, 1, "hard-swish", False, False, True),
    (2, 40, 10, 14, 120, 1, 1, "relu", False, False, True),
    (2, 16, 20, 20, 64, 1, 2, "relu6", False, False, True),     # 1x1 stride 2 (MobileNet expansion)
    (1, 200, 9, 9, 80, 1, 1, "linear", True, False, True),
    (2, 128, 33, 41, 256, 3, 1, "leaky", False, False, True),
    (4, 512, 16, 20, 1024, 3, 1, "leaky", False, False, True),
    (2, 2048, 8, 10, 512, 1, 1, "leaky", False, False, True),
    (1, 8, 64, 80, 32, 3, 1, "mish", False, False, True),
    (2, 64, 12, 12, 64, 5, 1, "relu", False, False, True),
    # early high-resolution 3x3 layers: CTA-pair kernel with the whole filter resident in shared memory
    (4, 32, 64, 80, 64, 3, 1, "leaky", True, False, True),
    (2, 64, 96, 80, 128, 3, 1, "mish", True, False, True),
    (3, 48, 70, 83, 96, 3, 1, "leaky", False, False, True),     # ragged tiles, 3 K steps, 96 output channels
    (5, 16, 50, 64, 64, 3, 1, "relu", False, False, True),
    # 1x1 layers with Cin * Cout <= 2048: CUDA-core kernel (conv_thin.cu), output slabs of 32 / 24 / 16 / 8 channels
    (2, 16, 33, 47, 16, 1, 1, "relu", True, False, True),
    (3, 72, 12, 14, 24, 1, 1, "linear", True, False, True),
    (2, 64, 9, 11, 24, 1, 1, "linear", False, False, True),
    (2, 8, 13, 9, 40, 1, 1, "relu", False, False, True),
    (2, 24, 21, 21, 72, 1, 2, "hard-swish", False, False, True),
    (2, 32, 10, 12, 64, 1, 1, "mish", True, False, True),
    (3, 40, 17, 19, 120, 1, 1, "relu", False, False, True),
    (2, 120, 11, 13, 40, 1, 1, "linear", True, False, True),
    (1, 64, 5, 7, 128, 1, 2, "leaky", False, False, True),
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_bn_act(case, dtype):
    from dyk import ops
    N, Cin, H, W, Cout, k, s, act, res, up, bn = case
    g = torch.Generator().manual_seed(hash(case) % 2 ** 31)
    conv = nn.Conv2d(Cin, Cout, k, s, k // 2, bias=not bn)
    bnm = nn.BatchNorm2d(Cout) if bn else None
    with torch.no_grad():
        conv.weight.copy_(_q(torch.randn(conv.weight.shape, generator=g) / (Cin * k * k) ** 0.5, dtype))
        if bn:
            bnm.weight.copy_(torch.rand(Cout, generator=g) + 0.5)
            bnm.bias.copy_(torch.randn(Cout, generator=g) * 0.2)
            bnm.running_mean.copy_(torch.randn(Cout, generator=g) * 0.2)
            bnm.running_var.copy_(torch.rand(Cout, generator=g) + 0.5)
        else:
            conv.bias.copy_(torch.randn(Cout, generator=g))
    x = _q(torch.randn((N, Cin, H, W), generator=g), dtype)
    conv, x = conv.to(DEV).eval(), x.to(DEV)
    bnm = bnm.to(DEV).eval() if bn else None
    with torch.no_grad():
        want = conv(x)
        if bn:
            want = bnm(want)
        want = _act_ref(want, act)
        r = None
        if res:
            r = _q(torch.randn(want.shape, generator=g), dtype).to(DEV)
            want = want + r
        if up:
            want = F.interpolate(want, scale_factor=2, mode="nearest")
    got = ops.conv_bn_act(x, conv, bnm, act, dtype=dtype, residual=r, upsample2x=up)
    assert got.shape == want.shape
    eps = 2e-3 if dtype == torch.float16 else 1.6e-2
    _close(got, want, rtol=eps, atol=eps, what=f"conv {case} {dtype}")


@pytest.mark.parametrize("in_dtype", [torch.float32, torch.uint8])
def test_stem_conv_reads_nchw_frames(in_dtype):
    from dyk import ops
    g = torch.Generator().manual_seed(3)
    conv = nn.Conv2d(3, 32, 3, 1, 1, bias=False)
    bn = nn.BatchNorm2d(32)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(32, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(32, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(32, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(32, generator=g) + 0.5)
    conv, bn = conv.to(DEV).eval(), bn.to(DEV).eval()
    if in_dtype == torch.uint8:
        x = torch.randint(0, 256, (2, 3, 40, 56), dtype=torch.uint8, generator=g).to(DEV)
        xf = x.float() / 255.0
    else:
        x = torch.rand((2, 3, 40, 56), generator=g).to(DEV)
        xf = x
    with torch.no_grad():
        want = F.leaky_relu(bn(conv(xf)), 0.1)
    got = ops.conv_bn_act(x, conv, bn, "leaky", dtype=torch.float16)
    _close(got, want, rtol=2e-3, atol=2e-3, what="stem")


@pytest.mark.parametrize("in_dtype", [torch.float32, torch.uint8])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("Cout,stride,H,W", [(16, 2, 40, 56), (16, 2, 37, 45), (16, 1, 21, 30), (32, 2, 33, 64), (32, 2, 18, 7)])
def test_stem_unrolled_kernel_equals_generic_kernel(Cout, stride, H, W, dt, in_dtype):
    """The unrolled 3x3 / Cin 3 stem (stem3x3_kernel: the MobileNet stems, models.py:35-36 with stride 2) against the generic
    CUDA-core stem it replaces (DYK_STEM_FAST=0): bit-identical (same fp32 order, byte -> v / 255 through a table of the same
    IEEE quotients, out-of-frame taps add exact zeros), and against fp32 torch."""
    import os
    from dyk import ops
    g = torch.Generator().manual_seed(Cout * 100 + H)
    N = 3
    if in_dtype == torch.uint8:
        x = torch.randint(0, 256, (N, 3, H, W), dtype=torch.uint8, generator=g).to(DEV)
        xf = x.float() / 255.0
    else:
        x = torch.rand((N, 3, H, W), generator=g).to(DEV)
        xf = x
    w = (torch.randn((Cout, 3, 3, 3), generator=g) * 0.3).to(DEV)              # OIHW
    sc = (torch.rand(Cout, generator=g) + 0.5).to(DEV)
    bi = (torch.randn(Cout, generator=g) * 0.1).to(DEV)
    Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    outs = []
    for mode in ("1", "0"):
        os.environ["DYK_STEM_FAST"] = mode
        try:
            y = ops.new_view(N, Ho, Wo, Cout, dt, DEV)
            ops.nhwc_stem(x, w.permute(0, 2, 3, 1).contiguous(), sc, bi, y, k=3, stride=stride, pad=1, act="hard-swish")
            torch.cuda.synchronize()
        finally:
            os.environ.pop("DYK_STEM_FAST", None)
        outs.append(y.buf)
    assert torch.equal(outs[0], outs[1]), "unrolled stem differs from the generic stem"
    want = F.hardswish(F.conv2d(xf, w, None, stride, 1) * sc.view(1, -1, 1, 1) + bi.view(1, -1, 1, 1))
    tol = 2e-3 if dt == torch.float16 else 1.6e-2
    _close(outs[0].permute(0, 3, 1, 2).float(), want, rtol=tol, atol=tol, what="stem3x3")


@pytest.mark.parametrize("k,stride,C", [(3, 1, 64), (3, 2, 72), (5, 1, 120), (5, 2, 40), (3, 1, 960)])
def test_depthwise_conv(k, stride, C):
    from dyk import ops
    g = torch.Generator().manual_seed(11)
    conv = nn.Conv2d(C, C, k, stride, k // 2, groups=C, bias=False)
    bn = nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(C, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(C, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    conv, bn = conv.to(DEV).eval(), bn.to(DEV).eval()
    x = _q(torch.randn((2, C, 13, 17), generator=g), torch.float16).to(DEV)
    with torch.no_grad():
        want = F.relu6(bn(conv(x)))
    got = ops.conv_bn_act(x, conv, bn, "relu6", dtype=torch.float16)
    _close(got, want, rtol=2e-3, atol=2e-3, what="dwconv")


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("k,stride,C,H,W,slack", [
    (3, 1, 16, 37, 50, 0), (3, 1, 72, 33, 41, 8), (3, 2, 64, 30, 45, 0), (3, 2, 128, 17, 20, 64), (5, 1, 120, 19, 23, 0),
    (5, 2, 40, 21, 37, 0), (5, 1, 960, 16, 20, 0), (3, 1, 184, 9, 11, 0), (5, 1, 200, 12, 10, 16), (3, 1, 296, 7, 9, 0),
    (5, 2, 672, 32, 40, 0), (3, 1, 8, 24, 70, 8), (3, 1, 480, 5, 3, 0)])
def test_depthwise_tile_kernel_equals_strip_kernel(dt, k, stride, C, H, W, slack):
    """The TMA-staged tile kernel (csrc/dwconv_tile.cu) against the strip kernel it replaces (bit-identical: same fp32
    accumulation order, zero padding adds exact zeros) and against fp32 torch (models.py:41 groups == C), on channel-slice
    views (pixel stride > C), ragged tiles, channel counts without a friendly block size and both strides."""
    import os
    from dyk import ops
    from dyk.ops import View
    g = torch.Generator().manual_seed(1000 * k + C + H)
    N = 3
    xb = torch.randn((N, H, W, C + slack), generator=g).to(dt).to(DEV)
    x = View(xb, slack, C)
    w = torch.randn((k, k, C), generator=g).to(DEV)
    sc = (torch.rand(C, generator=g) + 0.5).to(DEV)
    bi = (torch.randn(C, generator=g) * 0.1).to(DEV)
    Ho, Wo = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
    outs = []
    for mode in ("1", "0"):
        os.environ["DYK_DW_TILE"] = mode
        try:
            yb = torch.full((N, Ho, Wo, C + 8), 7.0, dtype=dt, device=DEV)
            ops.nhwc_dwconv(x, w, sc, bi, View(yb, 0, C), k=k, stride=stride, pad=k // 2, act="hard-swish")
            torch.cuda.synchronize()
        finally:
            os.environ.pop("DYK_DW_TILE", None)
        assert torch.all(yb[..., C:] == 7.0), "wrote outside the channel slice"
        outs.append(yb[..., :C])
    assert torch.equal(outs[0], outs[1]), "tile kernel differs from the strip kernel"
    xf = xb[..., slack:].float().permute(0, 3, 1, 2)
    want = F.hardswish(F.conv2d(xf, w.permute(2, 0, 1).unsqueeze(1), None, stride, k // 2, 1, C) * sc.view(1, -1, 1, 1)
                       + bi.view(1, -1, 1, 1))
    tol = 2e-3 if dt == torch.float16 else 1.6e-2
    _close(outs[0].permute(0, 3, 1, 2).float(), want, rtol=tol, atol=tol, what="dwconv tile")


def test_weighted_fusion_and_plain_shortcut():
    from dyk import ops
    g = torch.Generator().manual_seed(5)
    x = _q(torch.randn((2, 64, 9, 11), generator=g), torch.float16).to(DEV)
    a = _q(torch.randn((2, 64, 9, 11), generator=g), torch.float16).to(DEV)
    w = torch.tensor([0.3, -0.7], device=DEV)
    ww = torch.sigmoid(w) * (2 / 2)
    _close(ops.weighted_fusion(x, [a], w), x * ww[0] + a * ww[1], 2e-3, 2e-3, "weighted add")
    _close(ops.weighted_fusion(x, [a], None), x + a, 2e-3, 2e-3, "plain add")
    # channel mismatch rules (layers.py:78-83)
    a_small = a[:, :32].contiguous()
    want = x.clone()
    want[:, :32] = want[:, :32] + a_small
    _close(ops.weighted_fusion(x, [a_small], None), want, 2e-3, 2e-3, "slice input")
    x_small = x[:, :32].contiguous()
    _close(ops.weighted_fusion(x_small, [a], None), x_small + a[:, :32], 2e-3, 2e-3, "slice feature")
    wx = x * ww[0]
    wx[:, :32] = wx[:, :32] + a_small * ww[1]
    _close(ops.weighted_fusion(x, [a_small], w), wx, 2e-3, 2e-3, "weighted slice input")


@pytest.mark.parametrize("k", [3, 5, 9, 13])
def test_maxpool_same(k):
    from dyk import ops
    g = torch.Generator().manual_seed(k)
    x = _q(torch.randn((2, 64, 16, 20), generator=g), torch.float16).to(DEV)
    want = F.max_pool2d(x, k, 1, (k - 1) // 2)
    got = ops.maxpool(x, k, 1)
    assert torch.equal(got, want), "max-pool must be exact"


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(3, 72, 13, 19), (2, 512, 16, 20), (1, 8, 3, 2), (2, 40, 33, 41)])
def test_maxpool5_strip_kernel_exact(shape, dt):
    """The 5x5 / stride-1 SPP pool (maxpool5_kernel: packed 16-bit max over a sliding window, models.py:91-94) is exact for
    fp16 and bf16 on ragged widths, maps smaller than the window and channel-slice outputs."""
    from dyk import ops
    from dyk.ops import View
    g = torch.Generator().manual_seed(sum(shape))
    N, C, H, W = shape
    x = torch.randn(shape, generator=g).to(dt).to(DEV)
    want = F.max_pool2d(x.float(), 5, 1, 2).to(dt)
    xv = ops.to_nhwc(x, dt)
    yb = torch.full((N, H, W, C + 8), 3.0, dtype=dt, device=DEV)
    ops.nhwc_maxpool(xv, View(yb, 8, C), 5, 1)
    torch.cuda.synchronize()
    assert torch.all(yb[..., :8] == 3.0), "wrote outside the channel slice"
    assert torch.equal(yb[..., 8:].permute(0, 3, 1, 2), want), "max-pool must be exact"


def test_upsample_and_concat_exact():
    from dyk import ops
    g = torch.Generator().manual_seed(1)
    x = _q(torch.randn((2, 32, 5, 7), generator=g), torch.float16).to(DEV)
    y = _q(torch.randn((2, 16, 5, 7), generator=g), torch.float16).to(DEV)
    assert torch.equal(ops.upsample(x, 2), F.interpolate(x, scale_factor=2, mode="nearest"))
    assert torch.equal(ops.concat_channels([x, y, x]), torch.cat([x, y, x], 1))


@pytest.mark.parametrize("C,HW", [(256, (16, 20)), (72, (7, 9)), (1024, (4, 5)), (512, (64, 80))])
def test_squeeze_excitation(C, HW):
    from dyk import ops
    from build_utils.layers import SqueezeExcitation
    torch.manual_seed(C)
    se = SqueezeExcitation(C, 4).to(DEV)
    x = _q(torch.randn((2, C, *HW)), torch.float16).to(DEV)
    with torch.no_grad():
        s = F.adaptive_avg_pool2d(x, (1, 1))
        s = F.hardsigmoid(se.fc2(F.relu(se.fc1(s))))
        want = s * x
    got = ops.squeeze_excitation(x, se.fc1, se.fc2)
    _close(got, want, rtol=2e-3, atol=2e-3, what="SE")


@pytest.mark.parametrize("v4", [False, True])
def test_yolo_decode(v4):
    from models import YOLOLayer
    from oracle.darknet_ref import yolo_layer
    g = torch.Generator().manual_seed(9)
    anchors = np.array([[16., 42.], [22., 44.], [20., 53.]])
    p = (torch.randn((2, 18, 8, 10), generator=g) * 2).to(DEV)
    layer = YOLOLayer(anchors, 1, (64, 80), 8, "yolov4" if v4 else "yolov3").to(DEV).eval()
    io, pp = layer(p)
    io_ref, p_ref = yolo_layer(p.cpu(), torch.tensor(anchors, dtype=torch.float32), 8, 1, v4, False)
    assert torch.equal(pp.cpu(), p_ref)
    # sigmoid/exp are the device's libm vs the host's: a few ulp
    _close(io, io_ref, rtol=2e-6, atol=2e-6, what="decode")
    layer.train()
    assert torch.equal(layer(p).cpu(), p_ref)
    assert (layer.nx, layer.ny) == (10, 8)


def test_batched_weight_bind_matches_per_layer_ops():
    """dyk_fold_bn_multi / dyk_pack_weights_multi (one launch per model at bind time) against the per-layer torch fold
    of eval-mode BatchNorm2d (models.py:44-47) and the per-layer re-layout: scale / bias to 2 ulp, weights bit-equal."""
    from dyk import ops
    torch.manual_seed(5)
    layers = []
    for cin, cout, k, has_bn in ((16, 24, 3, True), (24, 18, 1, False), (40, 300, 3, True), (8, 8, 1, True)):
        conv = nn.Conv2d(cin, cout, k, bias=not has_bn).to(DEV)
        bn = nn.BatchNorm2d(cout, eps=1e-4).to(DEV) if has_bn else None
        if bn is not None:
            bn.weight.data.uniform_(0.5, 2.0); bn.bias.data.normal_()
            bn.running_mean.normal_(); bn.running_var.uniform_(0.01, 3.0)
        layers.append((conv, bn))
    outs = [ops.fold_bn_alloc(c, b) for c, b in layers]
    ops.fold_bn_multi([(c, b, s, t) for (c, b), (s, t) in zip(layers, outs)])
    for (c, b), (s, t) in zip(layers, outs):
        ws, wt = ops.fold_bn(c, b)
        assert (s is None) == (ws is None)
        if s is not None:
            torch.testing.assert_close(s, ws, rtol=3e-7, atol=0)
        torch.testing.assert_close(t, wt, rtol=3e-7, atol=1e-7)
    for dtype in (torch.float16, torch.bfloat16):
        bufs = [torch.empty((c.out_channels, c.kernel_size[0], c.kernel_size[0], c.in_channels), dtype=dtype, device=DEV)
                for c, _ in layers]
        ops.pack_conv_weights_multi([(c.weight.detach(), o) for (c, _), o in zip(layers, bufs)], dtype)
        for (c, _), o in zip(layers, bufs):
            assert torch.equal(o, ops.pack_conv_weight(c.weight, dtype))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 256, 16, 24, 256), (3, 128, 30, 41, 512), (1, 512, 16, 20, 320)],
                         ids=lambda s: "x".join(map(str, s)))
def test_weighted_fusion_fused_into_consumer_conv(shape, dtype):
    """WeightedFeatureFusion (layers.py:63-85) formed inside the consuming 3x3 convolution (dual-source operand of the
    CTA-pair halo kernel): bit-identical to the stand-alone add kernel followed by the same convolution, and within
    storage precision of conv(x * w0 + a * w1) in fp32."""
    from dyk import ops
    from dyk.ops import View
    N, Cin, H, W, Cout = shape
    g = torch.Generator().manual_seed(Cin + Cout)
    assert ops.conv_dual_source_supported(N, H, W, Cin, Cout, Cout, dtype, k=3, stride=1, pad=1)
    assert not ops.conv_dual_source_supported(N, H, W, Cin, Cout, Cout, dtype, k=3, stride=2, pad=1)
    assert not ops.conv_dual_source_supported(N, H, W, Cin, 64, 64, dtype, k=3, stride=1, pad=1)
    x = _q(torch.randn((N, Cin, H, W), generator=g), dtype).to(DEV)
    a = _q(torch.randn((N, Cin, H, W), generator=g), dtype).to(DEV)
    w_raw = torch.tensor([0.4, -1.1], device=DEV)
    weight = _q(torch.randn((Cout, Cin, 3, 3), generator=g) / (Cin * 9) ** 0.5, dtype).to(DEV)
    scale = torch.rand(1024, generator=g).to(DEV) + 0.5
    bias = torch.randn(1024, generator=g).to(DEV) * 0.1
    xv, av = ops.to_nhwc(x, dtype), ops.to_nhwc(a, dtype)
    wp = ops.pack_conv_weight(weight, dtype)
    # fused
    y1 = ops.new_view(N, H, W, Cout, dtype, DEV)
    ops.nhwc_conv(xv, wp, scale, bias, y1, k=3, stride=1, pad=1, act="leaky", x2=av, x_wts_raw=w_raw)
    # stand-alone add, then the same convolution
    wall = torch.empty(2, dtype=torch.float32, device=DEV)
    ops.fusion_weights(w_raw, wall)
    s = ops.new_view(N, H, W, Cin, dtype, DEV)
    ops.nhwc_add(xv, av, s, wall)
    y2 = ops.new_view(N, H, W, Cout, dtype, DEV)
    ops.nhwc_conv(s, wp, scale, bias, y2, k=3, stride=1, pad=1, act="leaky")
    torch.cuda.synchronize()
    assert torch.equal(y1.buf, y2.buf), "fused and stand-alone weighted fusion differ"
    ww = torch.sigmoid(w_raw)
    want = F.leaky_relu(F.conv2d((x * ww[0] + a * ww[1]).to(dtype).float(), weight, padding=1)
                        * scale[:Cout].view(1, -1, 1, 1) + bias[:Cout].view(1, -1, 1, 1), 0.1)
    eps = 2e-3 if dtype == torch.float16 else 1.6e-2
    _close(ops.to_nchw(y1), want, rtol=eps, atol=eps, what=f"dual-source conv {shape} {dtype}")


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(3, 72, 10, 14, 24), (2, 120, 16, 20, 40), (4, 672, 8, 10, 112), (2, 960, 9, 8, 160),
                                   (2, 1024, 16, 20, 512), (5, 64, 9, 9, 256)], ids=lambda s: "x".join(map(str, s)))
def test_se_gate_folded_into_consumer_conv(shape, dtype):
    """SqueezeExcitation's `scale * x` (layers.py:184-190) folded into the consuming 1x1 convolution as per-image weights:
    conv(x * g_n) = x . (W * g_n).  Bit-identical to running the same kernel image by image with the scaled weights, and
    within storage precision of conv(x * g) in fp32 (the reference's order)."""
    from dyk import ops
    from dyk.ops import View
    N, Cin, H, W, Cout = shape
    g = torch.Generator().manual_seed(Cin + 3 * Cout)
    assert ops.conv_gated_input_supported(H, W, Cin, k=1, stride=1, pad=0)
    assert not ops.conv_gated_input_supported(H, W, Cin, k=3, stride=1, pad=1)
    xf = _q(torch.randn((N, Cin, H, W), generator=g), dtype).to(DEV)
    x = ops.to_nhwc(xf, dtype)
    gate = torch.rand((N, Cin), generator=g).to(DEV)
    wf = _q(torch.randn((Cout, Cin, 1, 1), generator=g) / Cin ** 0.5, dtype).to(DEV)
    w = ops.pack_conv_weight(wf, dtype)
    scale = (torch.rand(1024, generator=g) + 0.5).to(DEV)
    bias = (torch.randn(1024, generator=g) * 0.1).to(DEV)
    wimg = torch.empty((N, Cout, 1, 1, Cin), dtype=dtype, device=DEV)
    ops.scale_weights_per_image(w, gate, wimg)
    assert torch.equal(wimg.view(N, Cout, Cin), (w.view(1, Cout, Cin).float() * gate.view(N, 1, Cin)).to(dtype))
    y1 = ops.new_view(N, H, W, Cout, dtype, DEV)
    ops.nhwc_conv(x, wimg, scale, bias, y1, k=1, stride=1, pad=0, act="hard-swish", cout=Cout, w_image_stride=Cout * Cin)
    import os
    os.environ["DYK_THIN"] = "0"             # the same (tcgen05) kernel, one image at a time, shared-weights path;
    try:                                     # without the switch the thin-layer kernel (conv_thin.cu) would take some shapes
        for n in range(N):
            xn = View(x.buf[n:n + 1], 0, Cin)
            yn = ops.new_view(1, H, W, Cout, dtype, DEV)
            ops.nhwc_conv(xn, wimg[n].contiguous(), scale, bias, yn, k=1, stride=1, pad=0, act="hard-swish")
            assert torch.equal(y1.buf[n], yn.buf[0]), f"image {n}: per-image-weight convolution differs"
    finally:
        os.environ.pop("DYK_THIN", None)
    want = F.hardswish(F.conv2d(xf * gate.view(N, Cin, 1, 1), wf) * scale[:Cout].view(1, -1, 1, 1) + bias[:Cout].view(1, -1, 1, 1))
    eps = 2e-3 if dtype == torch.float16 else 1.6e-2
    _close(ops.to_nchw(y1), want, rtol=eps, atol=eps, what=f"gate folded into weights {shape} {dtype}")


def test_se_gate_fused_into_weighted_shortcut():
    """dyk_fused_add_gated == dyk_scale_channels followed by dyk_fused_add, bit for bit (weighted and plain)."""
    from dyk import ops
    g = torch.Generator().manual_seed(13)
    N, C, H, W = 3, 96, 7, 9
    x = ops.to_nhwc(_q(torch.randn((N, C, H, W), generator=g), torch.float16).to(DEV), torch.float16)
    b = ops.to_nhwc(_q(torch.randn((N, C, H, W), generator=g), torch.float16).to(DEV), torch.float16)
    gate = torch.rand((N, C), generator=g).to(DEV)
    wall = torch.tensor([0.7, 1.2], device=DEV)
    xs = ops.new_view(N, H, W, C, torch.float16, DEV)
    ops.scale_channels(x, gate, xs)
    for wts in (wall, None):
        y1 = ops.new_view(N, H, W, C, torch.float16, DEV)
        y2 = ops.new_view(N, H, W, C, torch.float16, DEV)
        ops.nhwc_add(x, b, y1, wts, gate=gate)
        ops.nhwc_add(xs, b, y2, wts)
        assert torch.equal(y1.buf, y2.buf)
