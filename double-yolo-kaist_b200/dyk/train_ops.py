"""Launch wrappers for the training-side entry points of include/dyk_b200.h (train-mode BatchNorm, backward kernels,
tcgen05 weight gradient, data gradient through the forward conv kernels on rotated weights).

Like dyk/ops.py nothing here computes with PyTorch: torch owns memory and the stream.  Scratch buffers are cached per
(device, stream) and grown on demand.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _native as nat
from . import ops
from .ops import View, _p, _stream

_scratch = {}            # default registry (direct, eager use of the wrappers below)
_active = [_scratch]     # innermost registry in use; TrainPlan installs its own for the duration of a step


class use_scratch:
    """Context manager: scratch buffers requested inside belong to `registry` (a dict owned by the caller).

    A training plan captures its launches into CUDA graphs, which bake the scratch addresses in.  A buffer that a later,
    larger request replaces must therefore stay allocated for as long as the plan lives — with one process-wide registry
    a bigger plan (multi-scale training) or another model freed the old buffer, the allocator handed the memory to
    something else, and the next replay of the first plan's graph scribbled over it.  Each plan now owns its registry;
    replaced buffers are parked in registry['_keep'] and die with the plan."""

    def __init__(self, registry: dict):
        self.registry = registry

    def __enter__(self):
        _active.append(self.registry)

    def __exit__(self, *exc):
        _active.pop()


def scratch(nbytes: int, device, tag="f") -> torch.Tensor:
    """Device scratch of at least nbytes (uint8), cached per (device, stream, tag) in the active registry."""
    reg = _active[-1]
    key = (device, torch.cuda.current_stream().cuda_stream, tag)
    t = reg.get(key)
    if t is None or t.numel() < nbytes:
        if t is not None and reg is not _scratch:
            reg.setdefault("_keep", []).append(t)
        t = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        reg[key] = t
    return t


def _ws_floats(n, device, tag="f"):
    return C.c_void_p(scratch(4 * n, device, tag).data_ptr())


_counters = {}


def _tickets(device):
    """Zero-initialised ticket counters for the reductions that finalise in their last block (dyk_bn_train_stats /
    dyk_bn_act_bwd): one array per (device, stream) — the kernels leave it zero, consecutive launches on a stream share it.
    Opt-in (DYK_BN_FUSED_FINALIZE=1): measured on the dyolov4 bs-16 step the fused variant is 0.4 ms SLOWER than the
    separate finalisation launches inside a CUDA graph (43.2 vs 42.8 ms) — the reduction kernel cannot retire until its
    last block has walked the slab partials, which costs more than the ~2 us graph edge it saves — so the default is NULL."""
    if os.environ.get("DYK_BN_FUSED_FINALIZE", "0") != "1":
        return None
    key = (device, torch.cuda.current_stream().cuda_stream)
    t = _counters.get(key)
    if t is None:
        t = torch.zeros(8192, dtype=torch.int32, device=device)
        _counters[key] = t
    return C.c_void_p(t.data_ptr())


# ------------------------------------------------------------------------------------------ BatchNorm (training)
def bn_train_stats(z: View, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, mean, invstd) -> None:
    """Batch statistics of z -> scale/shift (y = z*scale + shift), saved mean/invstd, running stats updated in place."""
    Cc = z.C
    nat.call("dyk_bn_train_stats", z.ptr, z.stride, z.npix, Cc, z.dt, _p(gamma), _p(beta), float(eps), float(momentum),
             _p(running_mean), _p(running_var), _p(scale), _p(shift), _p(mean), _p(invstd),
             _ws_floats(nat.TRAIN_MAX_SLABS * 2 * Cc, z.buf.device), _tickets(z.buf.device), _stream())
    nat.count_launches(2)


def bn_act_apply(z: View, scale, shift, act: str, y: View) -> None:
    nat.call("dyk_bn_act_apply", z.ptr, z.stride, _p(scale), _p(shift), nat.ACT_IDS[act], y.ptr, y.stride, z.npix, z.C,
             z.dt, _stream())
    nat.count_launches()


def bn_act_bwd(dy: View, z: View, scale, shift, mean, invstd, gamma, act: str, dz: View, dgamma, dbeta) -> None:
    Cc = z.C
    nat.call("dyk_bn_act_bwd", dy.ptr, dy.stride, z.ptr, z.stride, _p(scale), _p(shift), _p(mean), _p(invstd), _p(gamma),
             nat.ACT_IDS[act], z.npix, Cc, z.dt, dz.ptr, dz.stride, _p(dgamma), _p(dbeta),
             _ws_floats((nat.TRAIN_MAX_SLABS * 2 + 3) * Cc, z.buf.device), _tickets(z.buf.device), _stream())
    nat.count_launches(3)


def chan_sum(x: View, out: torch.Tensor, accumulate: bool) -> None:
    nat.call("dyk_chan_sum", x.ptr, x.stride, x.npix, x.C, x.dt, _p(out), int(accumulate),
             _ws_floats(nat.TRAIN_MAX_SLABS * 2 * x.C, x.buf.device), _stream())
    nat.count_launches(2)


def axpby(src: View, dst: View, alpha: torch.Tensor = None, accumulate: bool = False) -> None:
    """dst = alpha*src (+ dst); alpha is a device fp32 scalar (None = 1)."""
    nat.call("dyk_axpby", src.ptr, src.stride, _p(alpha), dst.ptr, dst.stride, src.npix, src.C, int(accumulate), src.dt,
             _stream())
    nat.count_launches()


def fusion_weights_bwd(dy: View, a: View, b: View, w_raw: torch.Tensor, grad_w: torch.Tensor) -> None:
    nat.call("dyk_fusion_weights_bwd", dy.ptr, dy.stride, a.ptr, a.stride, b.ptr, b.stride, dy.npix, dy.C, dy.dt, _p(w_raw),
             w_raw.numel(), _p(grad_w), _ws_floats(2 * nat.TRAIN_MAX_SLABS * 2 * dy.C, dy.buf.device), _stream())
    nat.count_launches(3)


def maxpool_bwd(x: View, dy: View, dx: View, k: int, stride: int, accumulate: bool) -> None:
    pad = (k - 1) // 2
    Ho = (x.H + 2 * pad - k) // stride + 1
    Wo = (x.W + 2 * pad - k) // stride + 1
    idx = scratch(4 * x.N * Ho * Wo * x.C, x.buf.device, "i")
    nat.call("dyk_maxpool2d_bwd", x.ptr, x.stride, dy.ptr, dy.stride, dx.ptr, dx.stride, x.N, x.H, x.W, x.C, k, stride,
             int(accumulate), x.dt, C.c_void_p(idx.data_ptr()), _stream())
    nat.count_launches(2)


def upsample_bwd(dy: View, dx: View, s: int, accumulate: bool) -> None:
    nat.call("dyk_upsample_nearest_bwd", dy.ptr, dy.stride, dx.ptr, dx.stride, dx.N, dx.H, dx.W, dx.C, s, int(accumulate),
             dx.dt, _stream())
    nat.count_launches()


def se_bwd(x: View, dy: View, dx: View, w1, b1, w2, b2, pooled, gate, gw1, gb1, gw2, gb2, accumulate: bool) -> None:
    N, HW, Cc = x.N, x.H * x.W, x.C
    nat.call("dyk_se_bwd", x.ptr, x.stride, dy.ptr, dy.stride, dx.ptr, dx.stride, N, HW, Cc, _p(w1), _p(b1), _p(w2), _p(b2),
             w1.shape[0], _p(pooled), _p(gate), _p(gw1), _p(gb1), _p(gw2), _p(gb2), int(accumulate), x.dt,
             _ws_floats(72 * N * Cc, x.buf.device), _stream())
    nat.count_launches(N + 3)


def yolo_train_bwd(dp: torch.Tensor, dz: View) -> None:
    """dp fp32 (N, na, ny, nx, no) -> dz 16-bit NHWC (N, ny, nx, Cpad), zero padded channels."""
    N, na, ny, nx, no = dp.shape
    assert dz.c_off == 0 and dz.stride == dz.C
    nat.call("dyk_yolo_train_bwd", _p(dp), N, na, ny, nx, no, dz.ptr, dz.C, dz.dt, _stream())
    nat.count_launches()


# ------------------------------------------------------------------------------------------ convolution gradients
def pack_dgrad_weight(w_oihw: torch.Tensor, dtype, opad: int = None) -> torch.Tensor:
    """OIHW fp32 -> [I][kh][kw][Opad] 16-bit, taps rotated by 180 degrees (dyk_pack_weights_dgrad)."""
    O, I, kh, kw = w_oihw.shape
    opad = opad or O
    w = w_oihw.detach().to(torch.float32).contiguous()
    out = torch.empty((I, kh, kw, opad), dtype=dtype, device=w.device)
    nat.call("dyk_pack_weights_dgrad", _p(w), _p(out), O, I, kh, kw, opad, ops._DT[dtype], _stream())
    nat.count_launches()
    return out


def conv_dgrad(dz: View, w_dgrad: torch.Tensor, dx: View, *, k: int, stride: int, pad: int, accumulate: bool) -> None:
    """dx (+)= d(conv)/dx applied to dz.  w_dgrad from pack_dgrad_weight ([Cin][kh][kw][Cout_pad], Cout_pad == dz.C).

    stride 1: one forward convolution with the rotated weights and padding k-1-pad (residual operand = dx itself when
    accumulating).  stride 2: the four output parity planes of dx are four small stride-1 convolutions over dz (taps of
    matching parity), each stored through a parity-plane tensor map; accumulation then goes through a scratch tensor.
    """
    Cin = w_dgrad.shape[0]
    if stride == 1:
        ops.nhwc_conv(dz, w_dgrad, None, None, dx, k=k, stride=1, pad=k - 1 - pad, act="linear",
                      res=dx if accumulate else None, cout=Cin)
        return
    if stride != 2:
        raise nat.NativeError(f"conv_dgrad: stride {stride}")
    if dx.H != 2 * dz.H or dx.W != 2 * dz.W:
        raise nat.NativeError("conv_dgrad: stride-2 data gradient needs even input sizes (H, W multiples of 32 upstream)")
    target = dx
    if accumulate:
        target = ops.new_view(dx.N, dx.H, dx.W, dx.C, dx.buf.dtype, dx.buf.device)
    if k < 2:
        # 1x1 stride 2 (MobileNetV3 expand convs carry the stride): only one parity plane of dx receives a gradient;
        # the others are cleared (a memset of the gradient buffer, not arithmetic)
        target.buf[..., target.c_off:target.c_off + target.C].zero_()
    for ph in range(2):
        rows = [r for r in range(k) if (r - pad) % 2 == ph]          # taps that reach input rows of parity ph
        for pw in range(2):
            cols = [s for s in range(k) if (s - pad) % 2 == pw]
            if not rows or not cols:
                continue
            # dx[2i+ph] = sum_r dz[i + (ph + pad - r)/2] * W[r]; as a correlation over dz with offsets t = (ph+pad-r)/2 >= ...
            offs_h = sorted({(ph + pad - r) // 2 for r in rows})
            offs_w = sorted({(pw + pad - s) // 2 for s in cols})
            if offs_h[0] != offs_w[0]:
                raise nat.NativeError("conv_dgrad: asymmetric tap offsets are not supported")
            sub = _sub_dgrad_weights(w_dgrad, k, pad, ph, pw, offs_h, offs_w)
            p = nat.ConvParams()
            p.x, p.x_pix_stride = dz.ptr, dz.stride
            p.w = sub.data_ptr()
            p.y, p.y_pix_stride = target.ptr, target.stride
            p.N, p.H, p.W, p.Cin = dz.N, dz.H, dz.W, dz.C
            p.Cout = Cin
            p.Cout_store = target.C
            p.kh, p.kw = len(offs_h), len(offs_w)
            p.stride = 1
            p.pad = -offs_h[0]          # correlation out[m] = sum_u in[m + u - pad'] * K[u], first offset = -pad'
            p.act = nat.ACT_IDS["linear"]
            p.dtype = dz.dt
            p.out_h, p.out_w = dz.H, dz.W
            p.y_plane = 1 + ph * 2 + pw
            nat.call("dyk_conv2d_fwd", C.byref(p), _stream())
            nat.count_launches()
    if accumulate:
        axpby(target, dx, None, True)


def _sub_dgrad_weights(w_dgrad, k, pad, ph, pw, offs_h, offs_w):
    """Sub-kernel of the rotated dgrad weights for output parity (ph, pw): taps ordered by increasing dz offset.
    w_dgrad[ci][r'][s'][co] = W[co][ci][k-1-r'][k-1-s'];  dz offset t_h <-> original tap r = ph + pad - 2*t_h.
    (Not cached: w_dgrad is re-packed every step, and a pointer-keyed cache would alias recycled allocations.)"""
    rs = [k - 1 - (ph + pad - 2 * t) for t in offs_h]      # index into the rotated tensor: consecutive offsets = every 2nd tap
    ss = [k - 1 - (pw + pad - 2 * t) for t in offs_w]
    # strided slices, not index lists: no host->device index copy, so the step stays CUDA-graph capturable
    return w_dgrad[:, rs[0]:rs[-1] + 1:2, ss[0]:ss[-1] + 1:2].contiguous()   # layout-only gather of parameter taps


def conv_wgrad(x: View, dz: View, grad_w: torch.Tensor, *, k: int, stride: int, pad: int, accumulate: bool,
               cout_real: int = None, cin_real: int = None) -> None:
    """grad_w (fp32 OIHW) (+)= sum_p dz[p] (x) x[p + tap]  — tcgen05 split-K kernel + fixed-order reduction."""
    lib = nat.load()
    need = lib.dyk_conv2d_wgrad_workspace_bytes(x.C, dz.C, k)
    ws = scratch(need, x.buf.device, "w")
    nat.call("dyk_conv2d_wgrad", x.ptr, x.stride, dz.ptr, dz.stride, _p(grad_w), x.N, x.H, x.W, x.C, cin_real or x.C, dz.C,
             cout_real or dz.C, k, stride, pad, int(accumulate), x.dt, C.c_void_p(ws.data_ptr()), ws.numel(), _stream())
    nat.count_launches(2)


def dw_weight(conv) -> torch.Tensor:
    """state_dict [C][1][k][k] -> the depthwise kernels' fp32 [k][k][C] (layout only)."""
    Cc, _, k, _ = conv.weight.shape
    return conv.weight.detach().float().reshape(Cc, k, k).permute(1, 2, 0).contiguous()


def dwconv_dgrad(dz: View, w_kkc: torch.Tensor, dx: View, *, k: int, stride: int, pad: int, accumulate: bool) -> None:
    nat.call("dyk_dwconv2d_dgrad", dz.ptr, dz.stride, _p(w_kkc), dx.ptr, dx.stride, dx.N, dx.H, dx.W, dx.C, k, stride, pad,
             int(accumulate), dx.dt, _stream())
    nat.count_launches()


def dwconv_wgrad(x: View, dz: View, grad_w: torch.Tensor, *, k: int, stride: int, pad: int, accumulate: bool) -> None:
    nat.call("dyk_dwconv2d_wgrad", x.ptr, x.stride, dz.ptr, dz.stride, _p(grad_w), x.N, x.H, x.W, x.C, k, stride, pad,
             int(accumulate), x.dt, _ws_floats(nat.DW_WGRAD_SLABS * k * k * x.C, x.buf.device, "d"), _stream())
    nat.count_launches(2)


def stem_wgrad_tc(x_nchw: torch.Tensor, dz: View, grad_w: torch.Tensor, *, k: int, stride: int, pad: int,
                  accumulate: bool, resize: bool = False) -> None:
    """Stem weight gradient on the tensor cores: the frames are re-laid out once as NHWC with the channel dimension
    zero-padded to 8 (16-bit), then the generic split-K wgrad kernel runs with Cin = 8 and writes only the real
    input channels.  ~6x faster than the CUDA-core reduction below at 512x640 x 16."""
    N, Cin, H, W = x_nchw.shape
    kind = {torch.float32: 0, torch.uint8: 1}[x_nchw.dtype]
    dtype = dz.buf.dtype
    if resize and (H, W) != (dz.H, dz.W):
        # multi-scale step: the im2col rows are sampled from the bilinear resize of the original frames (dz has the resized size)
        if not (Cin == 3 and k == 3 and stride == 1 and pad == 1):
            raise nat.NativeError("stem_wgrad_tc: the fused resize exists for the 3x3 / stride-1 / 3-channel stem")
        Hd, Wd = dz.H, dz.W
        xc = scratch(2 * N * Hd * Wd * 32, x_nchw.device, "x32")[:2 * N * Hd * Wd * 32].view(dtype).view(N, Hd, Wd, 32)
        nat.call("dyk_frames_to_im2col32_resize", _p(x_nchw), _p(xc), N, H, W, Hd, Wd, dz.dt, kind, _stream())
        nat.count_launches()
        conv_wgrad(View(xc, 0, 32), dz, grad_w.view(grad_w.shape[0], 27, 1, 1), k=1, stride=1, pad=0, accumulate=accumulate,
                   cin_real=27)
        return
    if Cin == 3 and k == 3 and stride == 1 and pad == 1:
        # the 3x3 neighbourhood as 32 im2col "channels" (order ci, r, s = OIHW): a 1x1 weight gradient with one tap and
        # half-filled operand boxes instead of nine taps over boxes that are 7/8 zero padding (1.55 -> ~0.5 ms at 512x640x16)
        xc = scratch(2 * N * H * W * 32, x_nchw.device, "x32")[:2 * N * H * W * 32].view(dtype).view(N, H, W, 32)
        nat.call("dyk_frames_to_im2col32", _p(x_nchw), _p(xc), N, H, W, dz.dt, kind, _stream())
        nat.count_launches()
        conv_wgrad(View(xc, 0, 32), dz, grad_w.view(grad_w.shape[0], 27, 1, 1), k=1, stride=1, pad=0, accumulate=accumulate,
                   cin_real=27)
        return
    xp = scratch(2 * N * H * W * 8, x_nchw.device, "x8")[:2 * N * H * W * 8].view(dtype).view(N, H, W, 8)
    nat.call("dyk_frames_to_nhwc8", _p(x_nchw), _p(xp), N, Cin, H, W, dz.dt, kind, _stream())
    nat.count_launches()
    conv_wgrad(View(xp, 0, 8), dz, grad_w, k=k, stride=stride, pad=pad, accumulate=accumulate, cin_real=Cin)


def stem_wgrad(x_nchw: torch.Tensor, dz: View, grad_w: torch.Tensor, *, k: int, stride: int, pad: int, accumulate: bool) -> None:
    N, Cin, H, W = x_nchw.shape
    kind = {torch.float32: 0, torch.uint8: 1}[x_nchw.dtype]
    Cout = grad_w.shape[0]
    nat.call("dyk_conv2d_stem_wgrad", _p(x_nchw), dz.ptr, dz.stride, _p(grad_w), N, H, W, Cin, Cout, k, stride, pad,
             int(accumulate), dz.dt, kind, _ws_floats(nat.STEM_WGRAD_STRIPS * Cout * k * k * Cin, x_nchw.device, "s"), _stream())
    nat.count_launches(2)
