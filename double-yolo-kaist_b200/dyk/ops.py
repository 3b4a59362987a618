"""Thin launch wrappers over the C-ABI (include/dyk_b200.h) plus NCHW convenience entry points.

Two levels:
  * ``nhwc_*`` functions take :class:`View` objects (channel slices of channels-last buffers) and launch
    one native kernel on torch's current stream.  The execution plan (plan.py) is built from these.
  * the un-prefixed functions implement the module-level API of build_utils/layers.py for NCHW fp32
    CUDA tensors (convert in -> native kernels -> convert out).  They exist for API completeness and
    for the per-kernel parity tests; the model hot path never goes through them.

Nothing here computes with PyTorch: torch only owns the device memory and the stream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _native as nat

_DT = {torch.float16: nat.DYK_F16, torch.bfloat16: nat.DYK_BF16}
DEFAULT_DTYPE = torch.float16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise nat.NativeError(
            f"{what}: tensor is on {t.device}; the dual-stream YOLO hot path only exists as sm_100a "
            "kernels (no CPU fallback) — move the model and inputs to a B200")
    nat.check_device()


@dataclass
class View:
    """Channels [c_off, c_off + C) of a channels-last buffer ``buf`` of shape (N, H, W, Ctot)."""
    buf: torch.Tensor
    c_off: int
    C: int

    @property
    def N(self): return self.buf.shape[0]
    @property
    def H(self): return self.buf.shape[1]
    @property
    def W(self): return self.buf.shape[2]
    @property
    def stride(self): return self.buf.shape[3]
    @property
    def ptr(self): return self.buf.data_ptr() + self.c_off * self.buf.element_size()
    @property
    def dt(self): return _DT[self.buf.dtype]
    @property
    def npix(self): return self.N * self.H * self.W


def new_view(N, H, W, C, dtype, device) -> View:
    return View(torch.empty((N, H, W, C), dtype=dtype, device=device), 0, C)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


# ------------------------------------------------------------------------------------------ weights
def pack_conv_weight(w_oihw: torch.Tensor, dtype) -> torch.Tensor:
    """OIHW fp32 -> [O][kh][kw][I] in `dtype` (dyk_pack_weights_ohwi)."""
    O, I, kh, kw = w_oihw.shape
    w = w_oihw.detach().to(torch.float32).contiguous()
    out = torch.empty((O, kh, kw, I), dtype=dtype, device=w.device)
    nat.call("dyk_pack_weights_ohwi", _p(w), _p(out), O, I, kh, kw, _DT[dtype], _stream())
    nat.count_launches()
    return out


def pad_vec(v: torch.Tensor, mult: int = 256) -> torch.Tensor:
    n = v.numel()
    npad = (n + mult - 1) // mult * mult
    out = torch.zeros(npad, dtype=torch.float32, device=v.device)
    out[:n] = v.detach().to(torch.float32)
    return out


def fold_bn(conv, bn):
    """(scale, bias) fp32 vectors, padded to 256, equivalent to eval-mode BatchNorm2d after `conv`
    (reference models.py:47; eps/running stats semantics of nn.BatchNorm2d) or to the conv bias."""
    if bn is not None:
        inv = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
        g = bn.weight.detach().float() if bn.weight is not None else torch.ones_like(inv)
        b = bn.bias.detach().float() if bn.bias is not None else torch.zeros_like(inv)
        scale = g * inv
        bias = b - bn.running_mean.detach().float() * scale
        return pad_vec(scale), pad_vec(bias)
    if conv.bias is not None:
        return None, pad_vec(conv.bias)
    return None, None


def fold_bn_alloc(conv, bn, mult: int = 256):
    """Uninitialised (scale, bias) buffers of the shapes fold_bn returns; filled by fold_bn_multi."""
    dev = conv.weight.device
    npad = (conv.out_channels + mult - 1) // mult * mult
    if bn is not None:
        return (torch.empty(npad, dtype=torch.float32, device=dev), torch.empty(npad, dtype=torch.float32, device=dev))
    if conv.bias is not None:
        return None, torch.empty(npad, dtype=torch.float32, device=dev)
    return None, None


def _f32c(t, what):
    if t is None:
        return 0
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise nat.NativeError(f"{what} must be a contiguous float32 tensor")
    return t.data_ptr()


def fold_bn_multi(items) -> None:
    """items: (conv, bn, scale_out, bias_out) — fold_bn for all of them in one launch (dyk_fold_bn_multi)."""
    import struct
    rows = []
    for conv, bn, scale, bias in items:
        if scale is None and bias is None:
            continue
        Cc = conv.out_channels
        npad = (bias if bias is not None else scale).numel()
        if bn is not None:
            eps = struct.unpack("<i", struct.pack("<f", float(bn.eps)))[0]
            rows.append([_f32c(bn.weight, "BatchNorm weight"), _f32c(bn.bias, "BatchNorm bias"),
                         _f32c(bn.running_mean, "running_mean"), _f32c(bn.running_var, "running_var"), 0,
                         scale.data_ptr(), bias.data_ptr(), Cc, npad, eps])
        else:
            rows.append([0, 0, 0, 0, _f32c(conv.bias, "conv bias"), 0, bias.data_ptr(), Cc, npad, 0])
    if not rows:
        return
    dev = items[0][0].weight.device
    desc = torch.tensor(rows, dtype=torch.int64).to(dev)
    nat.call("dyk_fold_bn_multi", _p(desc), len(rows), _stream())
    nat.count_launches()


def pack_conv_weights_multi(items, dtype) -> None:
    """items: (weight OIHW fp32 parameter, packed [O][k][k][I] buffer) — all re-laid out by one dyk_pack_weights_multi."""
    rows, tiles = [], 0
    for w, out in items:
        O, I, kh, kw = w.shape
        if kh * kw > 9:
            raise nat.NativeError("pack_conv_weights_multi: kernels larger than 3x3 are packed one by one")
        rows.append([_f32c(w, "convolution weight"), out.data_ptr(), 0, O, I, kh * kw, 0, tiles])
        tiles += ((O + 31) // 32) * ((I + 31) // 32)
    if not rows:
        return
    desc = torch.tensor(rows, dtype=torch.int64).to(items[0][0].device)
    nat.call("dyk_pack_weights_multi", _p(desc), len(rows), tiles, _DT[dtype], _stream())
    nat.count_launches()


# ------------------------------------------------------------------------------------------ NHWC launches
def _conv_params(x: View, w_packed, scale, bias, y: View, *, k, stride, pad, act, res: View = None, upsample2x=False,
                 out_f32=False, cout=None, x2: View = None, x_wts_raw=None, w_image_stride=0, sm_limit=0):
    p = nat.ConvParams()
    p.w_image_stride = int(w_image_stride)
    p.sm_limit = int(sm_limit)
    p.x, p.x_pix_stride = x.ptr, x.stride
    p.w = None if w_packed is None else w_packed.data_ptr()
    p.scale = None if scale is None else scale.data_ptr()
    p.bias = None if bias is None else bias.data_ptr()
    p.y, p.y_pix_stride = y.ptr, y.stride
    if res is not None:
        p.res, p.res_pix_stride = res.ptr, res.stride
    if x2 is not None:
        if (x2.N, x2.H, x2.W, x2.C, x2.dt) != (x.N, x.H, x.W, x.C, x.dt):
            raise ValueError("dual-source convolution: the two inputs must have the same shape and dtype")
        p.x2, p.x2_pix_stride = x2.ptr, x2.stride
        p.x_wts_raw = x_wts_raw.data_ptr()
    p.N, p.H, p.W, p.Cin = x.N, x.H, x.W, x.C
    p.Cout = (w_packed.shape[0] if cout is None else cout)
    p.Cout_store = y.C
    p.kh = p.kw = k
    p.stride, p.pad = stride, pad
    p.act = nat.ACT_IDS[act]
    p.dtype = x.dt
    p.upsample2x = int(upsample2x)
    p.out_f32 = int(out_f32)
    return p


def nhwc_conv(x: View, w_packed: torch.Tensor, scale, bias, y: View, *, k, stride, pad, act, res: View = None,
              upsample2x=False, out_f32=False, cout=None, x2: View = None, x_wts_raw=None, w_image_stride=0,
              sm_limit=0) -> None:
    """x2 / x_wts_raw: the convolution's input is sigmoid(w)[0] * x + sigmoid(w)[1] * x2 (WeightedFeatureFusion fused into
    its consumer, build_utils/layers.py:63-85); see conv_dual_source_supported.
    w_image_stride != 0: w_packed holds one weight tensor per image, w_image_stride elements apart (a SqueezeExcitation
    gate folded into the consuming 1x1 convolution, scale_weights_per_image); pass cout explicitly."""
    p = _conv_params(x, w_packed, scale, bias, y, k=k, stride=stride, pad=pad, act=act, res=res, upsample2x=upsample2x,
                     out_f32=out_f32, cout=cout, x2=x2, x_wts_raw=x_wts_raw, w_image_stride=w_image_stride, sm_limit=sm_limit)
    nat.call("dyk_conv2d_fwd", C.byref(p), _stream())
    nat.count_launches()


def conv_gated_input_supported(H, W, Cin, *, k, stride, pad, upsample2x=False, out_f32=False) -> bool:
    """Can a SqueezeExcitation gate be folded into this convolution's weights (per-image weights, dyk_conv_params.
    w_image_stride)?  1x1 / stride 1 / pad 0 on maps large enough that one-image tiles of 128 pixels are not mostly padding."""
    return k == 1 and stride == 1 and pad == 0 and not upsample2x and Cin % 8 == 0 and H * W >= 64


def scale_weights_per_image(w_packed: torch.Tensor, gate: torch.Tensor, out: torch.Tensor) -> None:
    """out[n] = w_packed * gate[n] (1x1 convolution weights [Cout][1][1][Cin]; gate (N, Cin) fp32)."""
    Cout, Cin = w_packed.shape[0], w_packed.shape[-1]
    nat.call("dyk_scale_weights_per_image", _p(w_packed), _p(gate), gate.shape[-1], _p(out), gate.shape[0], Cout, Cin,
             _DT[w_packed.dtype], _stream())
    nat.count_launches()


def conv_dual_source_supported(N, H, W, Cin, cout, cout_store, dtype, *, k, stride, pad, upsample2x=False, out_f32=False) -> bool:
    """dyk_conv2d_dual_source_supported for a layer shape (no tensors needed)."""
    p = nat.ConvParams()
    p.N, p.H, p.W, p.Cin, p.Cout, p.Cout_store = N, H, W, Cin, cout, cout_store
    p.kh = p.kw = k
    p.stride, p.pad = stride, pad
    p.dtype = _DT[dtype]
    p.upsample2x, p.out_f32 = int(upsample2x), int(out_f32)
    return bool(nat.load().dyk_conv2d_dual_source_supported(C.byref(p)))


def nhwc_stem(x_nchw: torch.Tensor, w_f32_ohwi, scale, bias, y: View, *, k, stride, pad, act, resize_to=None) -> None:
    """resize_to=(H, W): the frames are read through a bilinear resize to H x W (F.interpolate(..., mode='bilinear',
    align_corners=False), kaist_train_eval_utils.py:59-71) — the resized batch is never materialised."""
    N, Cin, H, W = x_nchw.shape
    kind = {torch.float32: 0, torch.uint8: 1}[x_nchw.dtype]
    if resize_to is not None and tuple(resize_to) != (H, W):
        nat.call("dyk_conv2d_stem_nchw_resize_fwd", _p(x_nchw), _p(w_f32_ohwi), _p(scale), _p(bias), y.ptr, y.stride,
                 N, H, W, int(resize_to[0]), int(resize_to[1]), Cin, y.C, k, stride, pad, nat.ACT_IDS[act], y.dt, kind, _stream())
    else:
        nat.call("dyk_conv2d_stem_nchw_fwd", _p(x_nchw), _p(w_f32_ohwi), _p(scale), _p(bias), y.ptr, y.stride,
                 N, H, W, Cin, y.C, k, stride, pad, nat.ACT_IDS[act], y.dt, kind, _stream())
    nat.count_launches()


def nhwc_dwconv(x: View, w_kkc, scale, bias, y: View, *, k, stride, pad, act) -> None:
    nat.call("dyk_dwconv2d_fwd", x.ptr, x.stride, _p(w_kkc), _p(scale), _p(bias), y.ptr, y.stride, x.N, x.H, x.W,
             x.C, k, stride, pad, nat.ACT_IDS[act], x.dt, _stream())
    nat.count_launches()


def nhwc_add(a: View, b: View, y: View, wts=None, gate=None) -> None:
    """gate (N, C) fp32: operand a is seen through a SqueezeExcitation gate that was not materialised (dyk_fused_add_gated)."""
    if gate is not None:
        nat.call("dyk_fused_add_gated", a.ptr, a.stride, _p(gate), gate.shape[-1], a.H * a.W, b.ptr, b.stride, y.ptr, y.stride,
                 a.npix, y.C, _p(wts), a.dt, _stream())
    else:
        nat.call("dyk_fused_add", a.ptr, a.stride, b.ptr, b.stride, y.ptr, y.stride, a.npix, y.C, _p(wts), a.dt,
                 _stream())
    nat.count_launches()


def fusion_weights(w_raw: torch.Tensor, out: torch.Tensor) -> None:
    nat.call("dyk_fusion_weights", _p(w_raw), _p(out), w_raw.numel(), _stream())
    nat.count_launches()


def nhwc_copy(src: View, dst: View) -> None:
    nat.call("dyk_copy_slice", src.ptr, src.stride, dst.ptr, dst.stride, src.npix, src.C, src.dt, _stream())
    nat.count_launches()


def nhwc_maxpool(x: View, y: View, k, stride) -> None:
    nat.call("dyk_maxpool2d", x.ptr, x.stride, y.ptr, y.stride, x.N, x.H, x.W, x.C, k, stride, x.dt, _stream())
    nat.count_launches()


def nhwc_upsample(x: View, y: View, s) -> None:
    nat.call("dyk_upsample_nearest", x.ptr, x.stride, y.ptr, y.stride, x.N, x.H, x.W, x.C, s, x.dt, _stream())
    nat.count_launches()


def nhwc_se_gate(x: View, w1, b1, w2, b2, pooled, gate) -> None:
    """Only the gate of a SqueezeExcitation block (pool -> fc1 -> relu -> fc2 -> hardsigmoid, layers.py:184-189): gate[n, c];
    the `scale * x` is applied by the consumers (per-image weights of nhwc_conv / nhwc_add gate)."""
    nat.call("dyk_se_gate", x.ptr, x.stride, x.N, x.H * x.W, x.C, _p(w1), _p(b1), _p(w2), _p(b2), w1.shape[0],
             _p(pooled), _p(gate), x.dt, _stream())
    nat.count_launches(3)


def scale_channels(x: View, gate: torch.Tensor, y: View) -> None:
    """y = x * gate[n, c] (the last line of SqueezeExcitation.forward, layers.py:190)."""
    nat.call("dyk_scale_channels", x.ptr, x.stride, _p(gate), y.ptr, y.stride, x.N, x.H * x.W, x.C, x.dt, _stream())
    nat.count_launches()


def nhwc_se(x: View, y: View, w1, b1, w2, b2, pooled, gate) -> None:
    nat.call("dyk_se_gate", x.ptr, x.stride, x.N, x.H * x.W, x.C, _p(w1), _p(b1), _p(w2), _p(b2), w1.shape[0],
             _p(pooled), _p(gate), x.dt, _stream())
    nat.call("dyk_scale_channels", x.ptr, x.stride, _p(gate), y.ptr, y.stride, x.N, x.H * x.W, x.C, x.dt,
             _stream())
    nat.count_launches(4)


# ------------------------------------------------------------------------------------------ fp32-accurate mode
# (include/dyk_b200.h "fp32-accurate mode"; csrc/f32_path.cu).  Views are channel slices of fp32 NHWC buffers.
def f32_pack_conv_weight(w_oihw: torch.Tensor) -> torch.Tensor:
    """OIHW fp32 -> [O][kh][kw][6I] bf16 = [w1|w1|w1|w2|w2|w3], w = w1 + w2 + w3 (dyk_f32_pack_split6)."""
    O, I, kh, kw = w_oihw.shape
    w = w_oihw.detach().to(torch.float32).contiguous()
    out = torch.empty((O, kh, kw, 6 * I), dtype=torch.bfloat16, device=w.device)
    nat.call("dyk_f32_pack_split6", _p(w), _p(out), O, I, kh, kw, _stream())
    nat.count_launches()
    return out


def f32_conv(x: View, w_split: torch.Tensor, scale, bias, y: View, split_buf: torch.Tensor, *, k, stride, pad, act,
             cout) -> None:
    """act(scale * conv(x, w) + bias) with fp32 accuracy on the bf16 tensor-core kernel: x is split into the 6*Cin-channel
    bf16 operand [x1|x2|x3|x1|x2|x1] (scratch `split_buf`), the convolution accumulates the six partial products in fp32
    and its epilogue stores fp32."""
    n = x.npix * 6 * x.C
    if split_buf.numel() < n:
        raise nat.NativeError("f32_conv: split scratch too small")
    sv = View(split_buf[:n].view(x.N, x.H, x.W, 6 * x.C), 0, 6 * x.C)
    nat.call("dyk_f32_split6", x.ptr, x.stride, sv.ptr, x.npix, x.C, _stream())
    nat.count_launches()
    nhwc_conv(sv, w_split, scale, bias, y, k=k, stride=stride, pad=pad, act=act, out_f32=2, cout=cout)


def f32_stem(x_nchw: torch.Tensor, w_f32_ohwi, scale, bias, y: View, *, k, stride, pad, act) -> None:
    N, Cin, H, W = x_nchw.shape
    kind = {torch.float32: 0, torch.uint8: 1}[x_nchw.dtype]
    nat.call("dyk_f32_stem_nchw_fwd", _p(x_nchw), _p(w_f32_ohwi), _p(scale), _p(bias), y.ptr, y.stride, N, H, W, Cin, y.C, k,
             stride, pad, nat.ACT_IDS[act], kind, _stream())
    nat.count_launches()


def f32_dwconv(x: View, w_kkc, scale, bias, y: View, *, k, stride, pad, act) -> None:
    nat.call("dyk_f32_dwconv2d_fwd", x.ptr, x.stride, _p(w_kkc), _p(scale), _p(bias), y.ptr, y.stride, x.N, x.H, x.W, x.C, k,
             stride, pad, nat.ACT_IDS[act], _stream())
    nat.count_launches()


def f32_add(a: View, b: View, y: View, wts=None) -> None:
    nat.call("dyk_f32_fused_add", a.ptr, a.stride, None if b is None else b.ptr, 0 if b is None else b.stride, y.ptr,
             y.stride, a.npix, y.C, _p(wts), _stream())
    nat.count_launches()


def f32_copy(src: View, dst: View) -> None:
    f32_add(src, None, dst)


def f32_maxpool(x: View, y: View, k, stride) -> None:
    nat.call("dyk_f32_maxpool2d", x.ptr, x.stride, y.ptr, y.stride, x.N, x.H, x.W, x.C, k, stride, _stream())
    nat.count_launches()


def f32_upsample(x: View, y: View, s) -> None:
    nat.call("dyk_f32_upsample_nearest", x.ptr, x.stride, y.ptr, y.stride, x.N, x.H, x.W, x.C, s, _stream())
    nat.count_launches()


def f32_se(x: View, y: View, w1, b1, w2, b2, pooled, gate) -> None:
    nat.call("dyk_f32_se", x.ptr, x.stride, y.ptr, y.stride, x.N, x.H * x.W, x.C, _p(w1), _p(b1), _p(w2), _p(b2), w1.shape[0],
             _p(pooled), _p(gate), _stream())
    nat.count_launches(4)


def yolo_decode(p_head: torch.Tensor, p_stride, p_out, io_out, *, N, ny, nx, na, no, anchor_vec, stride, v4,
                rows_total, row_off, in_kind=2) -> None:
    nat.call("dyk_yolo_decode", _p(p_head), p_stride, _p(p_out), _p(io_out), N, ny, nx, na, no, _p(anchor_vec),
             float(stride), int(v4), rows_total, row_off, in_kind, _stream())
    nat.count_launches()


def to_nhwc(x_nchw: torch.Tensor, dtype=None) -> View:
    dtype = dtype or DEFAULT_DTYPE
    _require_cuda(x_nchw, "to_nhwc")
    x = x_nchw.detach().to(torch.float32).contiguous()
    N, Cc, H, W = x.shape
    v = new_view(N, H, W, Cc, dtype, x.device)
    nat.call("dyk_nchw_f32_to_nhwc", _p(x), v.ptr, v.stride, N, Cc, H, W, v.dt, _stream())
    nat.count_launches()
    return v


def to_nchw(v: View) -> torch.Tensor:
    out = torch.empty((v.N, v.C, v.H, v.W), dtype=torch.float32, device=v.buf.device)
    nat.call("dyk_nhwc_to_nchw_f32", v.ptr, v.stride, _p(out), v.N, v.C, v.H, v.W, v.dt, _stream())
    nat.count_launches()
    return out


# ------------------------------------------------------------------------------------------ NCHW module API
def _round8(c):
    return (c + 7) // 8 * 8


def conv_bn_act(x: torch.Tensor, conv, bn, act: str, dtype=None, residual: torch.Tensor = None,
                upsample2x=False) -> torch.Tensor:
    """act(BN_eval(conv(x))) (+ residual) for an NCHW CUDA tensor; dense, depthwise or stem convolution."""
    _require_cuda(x, "conv_bn_act")
    dtype = dtype or DEFAULT_DTYPE
    k = conv.kernel_size[0]
    stride = conv.stride[0]
    pad = conv.padding[0]
    Cin, Cout = conv.in_channels, conv.out_channels
    scale, bias = fold_bn(conv, bn)
    N, _, H, W = x.shape
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    if conv.groups == 1 and Cin <= 4:
        y = new_view(N, Ho, Wo, Cout, dtype, x.device)
        w = conv.weight.detach().float().permute(0, 2, 3, 1).contiguous()
        nhwc_stem(x.detach().contiguous(), w, scale, bias, y, k=k, stride=stride, pad=pad, act=act)
        return to_nchw(y)
    xv = to_nhwc(x, dtype)
    if conv.groups == 1:
        up = 2 if upsample2x else 1
        cs = _round8(Cout)
        y = new_view(N, Ho * up, Wo * up, cs, dtype, x.device)
        w = pack_conv_weight(conv.weight, dtype)
        rv = to_nhwc(residual, dtype) if residual is not None else None
        nhwc_conv(xv, w, scale, bias, y, k=k, stride=stride, pad=pad, act=act, res=rv, upsample2x=upsample2x)
        return to_nchw(View(y.buf, 0, Cout))
    if conv.groups == Cin and Cout == Cin:
        y = new_view(N, Ho, Wo, Cout, dtype, x.device)
        w = conv.weight.detach().float().reshape(Cout, k, k).permute(1, 2, 0).contiguous()
        nhwc_dwconv(xv, w, scale, bias, y, k=k, stride=stride, pad=pad, act=act)
        return to_nchw(y)
    raise nat.NativeError(f"conv_bn_act: groups={conv.groups} with Cin={Cin}, Cout={Cout} is not on the hot path")


def concat_channels(tensors, dtype=None) -> torch.Tensor:
    _require_cuda(tensors[0], "concat_channels")
    dtype = dtype or DEFAULT_DTYPE
    N, _, H, W = tensors[0].shape
    ctot = sum(t.shape[1] for t in tensors)
    out = new_view(N, H, W, ctot, dtype, tensors[0].device)
    off = 0
    for t in tensors:
        nhwc_copy(to_nhwc(t, dtype), View(out.buf, off, t.shape[1]))
        off += t.shape[1]
    return to_nchw(out)


def weighted_fusion(x: torch.Tensor, others, w_param, dtype=None) -> torch.Tensor:
    """WeightedFeatureFusion.forward (reference layers.py:63-85) incl. the channel-mismatch rules."""
    _require_cuda(x, "weighted_fusion")
    dtype = dtype or DEFAULT_DTYPE
    n = len(others) + 1
    xv = to_nhwc(x, dtype)
    wall = None
    if w_param is not None:
        wall = torch.empty(n, dtype=torch.float32, device=x.device)
        fusion_weights(w_param.detach().float().contiguous(), wall)
    one = torch.ones(1, dtype=torch.float32, device=x.device)
    for i, a in enumerate(others):
        av = to_nhwc(a, dtype)
        wts = None
        if wall is not None:
            # x carries w[0] on the first add only; later adds leave the running sum unscaled
            wts = torch.cat([wall[0:1] if i == 0 else one, wall[i + 1:i + 2]]).contiguous()
        cx, ca = xv.C, av.C
        c = min(cx, ca)
        if cx <= ca:  # same shape, or slice the feature: result has cx channels
            y = new_view(xv.N, xv.H, xv.W, cx, dtype, x.device)
            nhwc_add(View(xv.buf, xv.c_off, c), View(av.buf, av.c_off, c), y, wts)
        else:  # slice the input: only the first ca channels receive the addend
            y = new_view(xv.N, xv.H, xv.W, cx, dtype, x.device)
            if wts is not None and i == 0:
                # the remaining channels are still multiplied by w[0]
                zero = torch.zeros_like(xv.buf)
                w_only = torch.cat([wall[0:1], torch.zeros(1, dtype=torch.float32, device=x.device)]).contiguous()
                nhwc_add(xv, View(zero, 0, cx), y, w_only)
            else:
                nhwc_copy(xv, y)
            nhwc_add(View(xv.buf, xv.c_off, c), av, View(y.buf, 0, c), wts)
        xv = y
    return to_nchw(xv)


def maxpool(x: torch.Tensor, k: int, stride: int, dtype=None) -> torch.Tensor:
    _require_cuda(x, "maxpool")
    xv = to_nhwc(x, dtype)
    pad = (k - 1) // 2
    Ho = (xv.H + 2 * pad - k) // stride + 1
    Wo = (xv.W + 2 * pad - k) // stride + 1
    y = new_view(xv.N, Ho, Wo, xv.C, xv.buf.dtype, x.device)
    nhwc_maxpool(xv, y, k, stride)
    return to_nchw(y)


def upsample(x: torch.Tensor, s: int, dtype=None) -> torch.Tensor:
    _require_cuda(x, "upsample")
    xv = to_nhwc(x, dtype)
    y = new_view(xv.N, xv.H * s, xv.W * s, xv.C, xv.buf.dtype, x.device)
    nhwc_upsample(xv, y, s)
    return to_nchw(y)


def activation(x: torch.Tensor, act: str, dtype=None) -> torch.Tensor:
    """act(x) for an NCHW CUDA tensor through the BatchNorm-apply kernel with unit scale / zero shift."""
    _require_cuda(x, "activation")
    xv = to_nhwc(x, dtype)
    if xv.C % 8:
        raise nat.NativeError("activation: channel count must be a multiple of 8")
    y = new_view(xv.N, xv.H, xv.W, xv.C, xv.buf.dtype, x.device)
    one = torch.ones(xv.C, dtype=torch.float32, device=x.device)
    zero = torch.zeros(xv.C, dtype=torch.float32, device=x.device)
    nat.call("dyk_bn_act_apply", xv.ptr, xv.stride, _p(one), _p(zero), nat.ACT_IDS[act], y.ptr, y.stride, xv.npix, xv.C,
             xv.dt, _stream())
    nat.count_launches()
    return to_nchw(y)


def se_weights(fc1, fc2):
    w1 = fc1.weight.detach().float().reshape(fc1.out_channels, fc1.in_channels).contiguous()
    w2 = fc2.weight.detach().float().reshape(fc2.out_channels, fc2.in_channels).contiguous()
    return w1, fc1.bias.detach().float().contiguous(), w2, fc2.bias.detach().float().contiguous()


def squeeze_excitation(x: torch.Tensor, fc1, fc2, dtype=None) -> torch.Tensor:
    _require_cuda(x, "squeeze_excitation")
    xv = to_nhwc(x, dtype)
    y = new_view(xv.N, xv.H, xv.W, xv.C, xv.buf.dtype, x.device)
    w1, b1, w2, b2 = se_weights(fc1, fc2)
    pooled = torch.empty((xv.N, 32, xv.C), dtype=torch.float32, device=x.device)   # DYK_SE_MAX_SLABS partials
    gate = torch.empty((xv.N, xv.C), dtype=torch.float32, device=x.device)
    nhwc_se(xv, y, w1, b1, w2, b2, pooled, gate)
    return to_nchw(y)
