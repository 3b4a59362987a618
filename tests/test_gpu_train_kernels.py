"""Per-kernel parity of the training-side kernels on a B200 against plain PyTorch fp32 autograd of the same op
(the ops the reference's backward dispatches to).  Inputs are first rounded to the 16-bit storage dtype, so only
accumulation order and the final 16-bit rounding differ; tolerances are written per test."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DT = [torch.float16, torch.bfloat16]
EPS16 = {torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}


@pytest.fixture(scope="module", autouse=True)
def _setup(native_lib):
    assert torch.cuda.is_available(), "GPU tests need a B200"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from dyk import _native
    _native.check_device()


def _nhwc(t_nchw, dtype):
    """NCHW fp32 -> (View over an NHWC 16-bit copy, the rounded values as NCHW fp32)."""
    from dyk.ops import View
    q = t_nchw.to(dtype)
    buf = q.permute(0, 2, 3, 1).contiguous().to(DEV)
    return View(buf, 0, buf.shape[3]), q.float()


def _back(view):
    return view.buf[..., view.c_off:view.c_off + view.C].permute(0, 3, 1, 2).float().cpu()


def _new(N, H, W, Cc, dtype):
    from dyk import ops
    return ops.new_view(N, H, W, Cc, dtype, torch.device(DEV))


def _act(x, act):
    return {"linear": lambda v: v, "leaky": lambda v: F.leaky_relu(v, 0.1), "mish": F.mish, "relu": F.relu,
            "relu6": F.relu6, "hard-swish": F.hardswish, "hard-sigmoid": F.hardsigmoid}[act](x)


def _assert_close(got, want, rel, what, floor=None):
    err = (got - want).abs()
    scale = want.abs().clamp_min(floor if floor is not None else float(want.pow(2).mean().sqrt()) * 0.25)
    bad = err > rel * scale
    assert not bool(bad.any()), (what, "max rel err", float((err / scale).max()), int(bad.sum()), "of", bad.numel())


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("act", ["leaky", "mish", "linear", "relu6", "hard-swish"])
def test_bn_train_forward_and_backward(dtype, act):
    from dyk import train_ops as T
    g = torch.Generator().manual_seed(1)
    N, Cc, H, W = 3, 72, 13, 17
    z = torch.randn((N, Cc, H, W), generator=g) * 1.7 + 0.3
    dy = torch.randn((N, Cc, H, W), generator=g)
    gamma = torch.rand(Cc, generator=g) + 0.5
    beta = torch.randn(Cc, generator=g) * 0.2
    rm0, rv0 = torch.randn(Cc, generator=g) * 0.1, torch.rand(Cc, generator=g) + 0.5
    zv, zq = _nhwc(z, dtype)
    dyv, dyq = _nhwc(dy, dtype)
    f32 = lambda t: t.clone().to(DEV)
    rm, rv = f32(rm0), f32(rv0)
    scale, shift, mean, invstd = (torch.empty(Cc, device=DEV) for _ in range(4))
    T.bn_train_stats(zv, f32(gamma), f32(beta), 1e-5, 0.1, rm, rv, scale, shift, mean, invstd)
    yv = _new(N, H, W, Cc, dtype)
    T.bn_act_apply(zv, scale, shift, act, yv)
    # reference: fp32 batch norm in training mode on the rounded z
    zr = zq.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rmr, rvr = rm0.clone(), rv0.clone()
    yr = _act(F.batch_norm(zr, rmr, rvr, gr, br, True, 0.1, 1e-5), act)
    _assert_close(_back(yv), yr.detach(), 3 * EPS16[dtype], "bn+act forward")
    assert torch.allclose(rm.cpu(), rmr, rtol=1e-5, atol=1e-6) and torch.allclose(rv.cpu(), rvr, rtol=1e-5, atol=1e-6)
    # backward
    yr.backward(dyq)
    dzv = _new(N, H, W, Cc, dtype)
    dgamma, dbeta = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    T.bn_act_bwd(dyv, zv, scale, shift, mean, invstd, f32(gamma), act, dzv, dgamma, dbeta)
    _assert_close(_back(dzv), zr.grad, 4 * EPS16[dtype], "bn+act backward dz")
    _assert_close(dgamma.cpu(), gr.grad, 2e-4, "dgamma", floor=float(gr.grad.abs().mean()))
    _assert_close(dbeta.cpu(), br.grad, 2e-4, "dbeta", floor=float(br.grad.abs().mean()))
    # accumulation into the parameter gradients
    T.bn_act_bwd(dyv, zv, scale, shift, mean, invstd, f32(gamma), act, dzv, dgamma, dbeta)
    _assert_close(dgamma.cpu(), 2 * gr.grad, 2e-4, "dgamma accumulated", floor=float(gr.grad.abs().mean()))


WGRAD_CASES = [
    # (N, Cin, H, W, Cout, k, stride, pad)
    (2, 64, 16, 20, 128, 3, 1, 1),
    (2, 128, 16, 24, 128, 3, 1, 1),
    (3, 256, 8, 10, 512, 3, 1, 1),
    (2, 256, 16, 20, 128, 1, 1, 0),
    (2, 512, 8, 10, 256, 1, 1, 0),
    (2, 64, 32, 40, 128, 3, 2, 1),
    (2, 32, 33, 41, 64, 3, 1, 1),          # ragged spatial size
    (1, 1024, 4, 5, 24, 1, 1, 0),          # detection head, Cout padded 18 -> 24
    (2, 96, 12, 12, 40, 1, 1, 0),          # odd channel counts
    (4, 128, 64, 80, 256, 3, 1, 1),        # many pixel blocks -> several splits
    (2, 24, 16, 20, 72, 1, 2, 0),          # MobileNetV3 expand conv that carries the stride
    (2, 16, 16, 20, 48, 1, 1, 0),
    (2, 72, 16, 20, 24, 1, 1, 0),
]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("case", WGRAD_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_wgrad(dtype, case):
    from dyk import train_ops as T
    N, Cin, H, W, Cout, k, stride, pad = case
    g = torch.Generator().manual_seed(2)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    x = torch.randn((N, Cin, H, W), generator=g)
    dz = torch.randn((N, Cout, Ho, Wo), generator=g)
    cout_real = 18 if Cout == 24 else Cout
    if cout_real != Cout:
        dz[:, cout_real:] = 0
    xv, xq = _nhwc(x, dtype)
    dzv, dzq = _nhwc(dz, dtype)
    grad = torch.full((cout_real, Cin, k, k), 7.0, device=DEV)   # must be overwritten
    T.conv_wgrad(xv, dzv, grad, k=k, stride=stride, pad=pad, accumulate=False, cout_real=cout_real)
    want = torch.nn.grad.conv2d_weight(xq.double(), (cout_real, Cin, k, k), dzq[:, :cout_real].double(), stride=stride,
                                       padding=pad).float()
    # fp32 accumulation of N*Ho*Wo products of O(1) values: error ~ 1e-6 * sqrt(K) relative to the rms
    _assert_close(grad.cpu(), want, 2e-4, ("wgrad", case), floor=float(want.pow(2).mean().sqrt()))
    T.conv_wgrad(xv, dzv, grad, k=k, stride=stride, pad=pad, accumulate=True, cout_real=cout_real)
    _assert_close(grad.cpu(), 2 * want, 2e-4, ("wgrad accumulate", case), floor=float(want.pow(2).mean().sqrt()))


DGRAD_CASES = [
    (2, 64, 16, 20, 128, 3, 1, 1),
    (2, 256, 8, 10, 512, 3, 1, 1),
    (2, 256, 16, 20, 128, 1, 1, 0),
    (2, 64, 32, 40, 128, 3, 2, 1),
    (1, 128, 16, 16, 256, 3, 2, 1),
    (1, 1024, 4, 5, 32, 1, 1, 0),           # head: Cout 18 padded to 32 for the dgrad GEMM's K
    (2, 24, 16, 20, 72, 1, 2, 0),           # 1x1 stride 2: three of the four parity planes of dx are zero
    (2, 16, 16, 20, 48, 1, 1, 0),
    (2, 72, 16, 20, 24, 1, 1, 0),
]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("case", DGRAD_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_dgrad(dtype, case):
    from dyk import train_ops as T
    N, Cin, H, W, Cout, k, stride, pad = case
    g = torch.Generator().manual_seed(3)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    cout_real = 18 if Cout == 32 else Cout
    w = torch.randn((cout_real, Cin, k, k), generator=g) / (Cin * k * k) ** 0.5
    dz = torch.randn((N, Cout, Ho, Wo), generator=g)
    dz[:, cout_real:] = 0
    dzv, dzq = _nhwc(dz, dtype)
    wq = w.to(dtype).float()
    wd = T.pack_dgrad_weight(w.to(DEV), dtype, opad=Cout)
    dx = _new(N, H, W, Cin, dtype)
    T.conv_dgrad(dzv, wd, dx, k=k, stride=stride, pad=pad, accumulate=False)
    want = torch.nn.grad.conv2d_input((N, Cin, H, W), wq.double(), dzq[:, :cout_real].double(), stride=stride,
                                      padding=pad).float()
    _assert_close(_back(dx), want, 3 * EPS16[dtype], ("dgrad", case))
    T.conv_dgrad(dzv, wd, dx, k=k, stride=stride, pad=pad, accumulate=True)
    _assert_close(_back(dx), 2 * want, 6 * EPS16[dtype], ("dgrad accumulate", case))


DW_CASES = [  # N, C, H, W, k, stride, pad   (pad 1 with k 5: DepthwiseSeparableConv2d's fixed padding, layers.py:224)
    (2, 72, 17, 23, 3, 1, 1), (2, 48, 16, 20, 3, 2, 1), (3, 120, 13, 11, 5, 1, 2), (2, 72, 16, 24, 5, 2, 2),
    (1, 960, 8, 10, 5, 1, 2), (2, 64, 9, 14, 5, 1, 1), (4, 16, 64, 80, 3, 1, 1),
]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("case", DW_CASES, ids=lambda c: "x".join(map(str, c)))
def test_dwconv_backward(dtype, case):
    """Depthwise data and weight gradients against autograd's conv2d_input / conv2d_weight (groups = C) in fp64."""
    from dyk import train_ops as T
    from dyk.ops import _p
    N, Cc, H, W, k, stride, pad = case
    g = torch.Generator().manual_seed(8)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    w = torch.randn((Cc, 1, k, k), generator=g) / k
    x = torch.randn((N, Cc, H, W), generator=g)
    dz = torch.randn((N, Cc, Ho, Wo), generator=g)
    xv, xq = _nhwc(x, dtype)
    dzv, dzq = _nhwc(dz, dtype)
    wk = w.reshape(Cc, k, k).permute(1, 2, 0).contiguous().to(DEV)
    dx = _new(N, H, W, Cc, dtype)
    T.dwconv_dgrad(dzv, wk, dx, k=k, stride=stride, pad=pad, accumulate=False)
    want = torch.nn.grad.conv2d_input((N, Cc, H, W), w.double(), dzq.double(), stride=stride, padding=pad, groups=Cc).float()
    _assert_close(_back(dx), want, 3 * EPS16[dtype], ("dw dgrad", case))
    T.dwconv_dgrad(dzv, wk, dx, k=k, stride=stride, pad=pad, accumulate=True)
    _assert_close(_back(dx), 2 * want, 6 * EPS16[dtype], ("dw dgrad accumulate", case))
    gw = torch.full((Cc, 1, k, k), 7.0, device=DEV)
    T.dwconv_wgrad(xv, dzv, gw, k=k, stride=stride, pad=pad, accumulate=False)
    want_w = torch.nn.grad.conv2d_weight(xq.double(), (Cc, 1, k, k), dzq.double(), stride=stride, padding=pad, groups=Cc).float()
    _assert_close(gw.cpu(), want_w, 2e-5, ("dw wgrad", case))
    first = gw.clone()
    T.dwconv_wgrad(xv, dzv, gw, k=k, stride=stride, pad=pad, accumulate=True)
    assert torch.equal(gw, 2 * first), "accumulating the same gradient twice must double it exactly (fixed reduction order)"


@pytest.mark.parametrize("dtype", DT)
def test_stem_wgrad(dtype):
    from dyk import train_ops as T
    g = torch.Generator().manual_seed(4)
    N, H, W, Cout = 2, 64, 96, 32
    x8 = torch.randint(0, 256, (N, 3, H, W), dtype=torch.uint8, generator=g)
    dz = torch.randn((N, Cout, H, W), generator=g)
    dzv, dzq = _nhwc(dz, dtype)
    want = torch.nn.grad.conv2d_weight((x8.float() / 255.0).double(), (Cout, 3, 3, 3), dzq.double(), stride=1, padding=1).float()
    for xin in (x8.to(DEV), (x8.float() / 255.0).to(DEV)):
        grad = torch.empty((Cout, 3, 3, 3), device=DEV)
        T.stem_wgrad(xin, dzv, grad, k=3, stride=1, pad=1, accumulate=False)
        _assert_close(grad.cpu(), want, 2e-4, "stem wgrad", floor=float(want.pow(2).mean().sqrt()))
        # tensor-core variant: the frames are rounded to the 16-bit type first (one rounding per pixel, averaged out)
        grad2 = torch.empty((Cout, 3, 3, 3), device=DEV)
        T.stem_wgrad_tc(xin, dzv, grad2, k=3, stride=1, pad=1, accumulate=False)
        _assert_close(grad2.cpu(), want, 4 * EPS16[dtype], "stem wgrad (tensor core)", floor=float(want.pow(2).mean().sqrt()))


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("k", [5, 9, 13, 3])
def test_maxpool_bwd(dtype, k):
    from dyk import train_ops as T
    g = torch.Generator().manual_seed(5)
    N, Cc, H, W = 2, 40, 16, 20
    x = torch.randn((N, Cc, H, W), generator=g)
    dy = torch.randn((N, Cc, H, W), generator=g)
    xv, xq = _nhwc(x, dtype)      # 16-bit values: ties between distinct pixels do occur
    dyv, dyq = _nhwc(dy, dtype)
    xr = xq.clone().requires_grad_(True)
    F.max_pool2d(xr, k, 1, (k - 1) // 2).backward(dyq)
    dx = _new(N, H, W, Cc, dtype)
    T.maxpool_bwd(xv, dyv, dx, k, 1, accumulate=False)
    _assert_close(_back(dx), xr.grad, 3 * EPS16[dtype], ("maxpool bwd", k))
    T.maxpool_bwd(xv, dyv, dx, k, 1, accumulate=True)
    _assert_close(_back(dx), 2 * xr.grad, 6 * EPS16[dtype], ("maxpool bwd accumulate", k))


@pytest.mark.parametrize("dtype", DT)
def test_upsample_axpby_chan_sum_yolo_bwd(dtype):
    from dyk import train_ops as T
    g = torch.Generator().manual_seed(6)
    N, Cc, H, W = 2, 48, 7, 9
    dy = torch.randn((N, Cc, 2 * H, 2 * W), generator=g)
    dyv, dyq = _nhwc(dy, dtype)
    dx = _new(N, H, W, Cc, dtype)
    T.upsample_bwd(dyv, dx, 2, accumulate=False)
    want = dyq.view(N, Cc, H, 2, W, 2).sum((3, 5))
    _assert_close(_back(dx), want, 3 * EPS16[dtype], "upsample bwd")
    # axpby with a device scalar, then accumulate
    alpha = torch.tensor([0.37], device=DEV)
    out = _new(N, 2 * H, 2 * W, Cc, dtype)
    T.axpby(dyv, out, alpha, accumulate=False)
    _assert_close(_back(out), 0.37 * dyq, 3 * EPS16[dtype], "axpby")
    first = _back(out)
    T.axpby(dyv, out, None, accumulate=True)
    _assert_close(_back(out), first + dyq, 3 * EPS16[dtype], "axpby accumulate")
    # per-channel sum (bias gradient)
    s = torch.zeros(Cc, device=DEV)
    T.chan_sum(dyv, s, accumulate=False)
    _assert_close(s.cpu(), dyq.sum((0, 2, 3)), 1e-4, "chan_sum", floor=float(dyq.sum((0, 2, 3)).abs().mean()))
    # yolo permute backward with channel padding
    dp = torch.randn((N, 3, H, W, 6), generator=g)
    dz = _new(N, H, W, 24, dtype)
    T.yolo_train_bwd(dp.to(DEV), dz)
    want = torch.zeros((N, 24, H, W))
    want[:, :18] = dp.permute(0, 1, 4, 2, 3).reshape(N, 18, H, W)
    _assert_close(_back(dz), want.to(dtype).float(), 1e-6, "yolo train bwd", floor=1.0)


@pytest.mark.parametrize("dtype", DT)
def test_fusion_weights_bwd(dtype):
    from dyk import train_ops as T
    g = torch.Generator().manual_seed(7)
    N, Cc, H, W = 2, 64, 12, 10
    a, b, dy = (torch.randn((N, Cc, H, W), generator=g) for _ in range(3))
    av, aq = _nhwc(a, dtype)
    bv, bq = _nhwc(b, dtype)
    dyv, dyq = _nhwc(dy, dtype)
    w = torch.tensor([0.3, -0.8], requires_grad=True)
    ww = torch.sigmoid(w) * (2 / 2)
    (aq * ww[0] + bq * ww[1]).backward(dyq)
    grad = torch.zeros(2, device=DEV)
    T.fusion_weights_bwd(dyv, av, bv, w.detach().to(DEV), grad)
    _assert_close(grad.cpu(), w.grad, 1e-4, "fusion w grad", floor=float(w.grad.abs().min()))


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("shape", [(2, 64, 16, 20), (3, 256, 9, 7), (2, 1024, 16, 20)])
def test_se_bwd(dtype, shape):
    from dyk import ops
    from dyk import train_ops as T
    import torch.nn as nn
    g = torch.Generator().manual_seed(8)
    N, Cc, H, W = shape
    Csq = max(8, (Cc // 4 + 7) // 8 * 8)
    x = torch.randn((N, Cc, H, W), generator=g)
    dy = torch.randn((N, Cc, H, W), generator=g)
    fc1, fc2 = nn.Conv2d(Cc, Csq, 1), nn.Conv2d(Csq, Cc, 1)
    with torch.no_grad():
        fc2.bias.uniform_(-2.5, 2.5, generator=g)     # exercise both regimes of the hard-sigmoid
    xv, xq = _nhwc(x, dtype)
    dyv, dyq = _nhwc(dy, dtype)
    xr = xq.clone().requires_grad_(True)
    s = F.adaptive_avg_pool2d(xr, 1)
    s = F.hardsigmoid(fc2(F.relu(fc1(s))))
    (s * xr).backward(dyq)
    w1, b1, w2, b2 = (t.to(DEV) for t in ops.se_weights(fc1, fc2))
    pooled = torch.empty((N, 32, Cc), device=DEV)
    gate = torch.empty((N, Cc), device=DEV)
    yv = _new(N, H, W, Cc, dtype)
    ops.nhwc_se(xv, yv, w1, b1, w2, b2, pooled, gate)
    gw1, gb1, gw2, gb2 = (torch.zeros_like(t) for t in (w1, b1, w2, b2))
    dx = _new(N, H, W, Cc, dtype)
    T.se_bwd(xv, dyv, dx, w1, b1, w2, b2, pooled, gate, gw1, gb1, gw2, gb2, accumulate=False)
    _assert_close(_back(dx), xr.grad, 4 * EPS16[dtype], "se dx")
    for got, want, nm in ((gw1, fc1.weight.grad.view_as(gw1.cpu()), "gw1"), (gb1, fc1.bias.grad, "gb1"),
                          (gw2, fc2.weight.grad.view_as(gw2.cpu()), "gw2"), (gb2, fc2.bias.grad, "gb2")):
        _assert_close(got.cpu(), want, 5e-4, ("se", nm), floor=float(want.abs().mean()) + 1e-8)
