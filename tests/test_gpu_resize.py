"""Multi-scale bilinear resize fused into the stem kernels on a B200.

The reference's training loop resizes the batch with `F.interpolate(imgs, size=ns, mode='bilinear', align_corners=False)`
(train_utils/kaist_train_eval_utils.py:59-71) before the model; `model(v, l, input_size=ns)` takes the ORIGINAL frames and
the stem convolution / the stem weight-gradient operand sample the resized frame on the fly.  Checked against PyTorch's own
interpolate (the op the reference calls): per kernel to 16-bit rounding, and end to end against the same native model fed
with pre-resized frames (eval outputs, training loss and parameter gradients)."""
from pathlib import Path

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TINY = str(Path(__file__).resolve().parent / "data" / "tiny_yolov3_train.cfg")


def _interp(x, size):
    x = x.float() / 255.0 if x.dtype == torch.uint8 else x
    return F.interpolate(x, size=size, mode="bilinear", align_corners=False)


@pytest.mark.parametrize("in_dtype", [torch.uint8, torch.float32])
@pytest.mark.parametrize("src,dst", [((40, 56), (64, 96)), ((72, 88), (48, 64)), ((50, 70), (50, 70)), ((33, 47), (96, 160))])
def test_stem_samples_the_resized_frame(native_lib, in_dtype, src, dst):
    from dyk import ops
    g = torch.Generator().manual_seed(7)
    conv = nn.Conv2d(3, 32, 3, 1, 1, bias=False)
    bn = nn.BatchNorm2d(32)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(32, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(32, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(32, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(32, generator=g) + 0.5)
    conv, bn = conv.to(DEV).eval(), bn.to(DEV).eval()
    if in_dtype == torch.uint8:
        x = torch.randint(0, 256, (2, 3) + src, dtype=torch.uint8, generator=g).to(DEV)
    else:
        x = torch.rand((2, 3) + src, generator=g).to(DEV)
    with torch.no_grad():
        want = F.leaky_relu(bn(conv(_interp(x, dst))), 0.1)
    scale, bias = ops.fold_bn(conv, bn)
    y = ops.new_view(2, dst[0], dst[1], 32, torch.float16, DEV)
    w = conv.weight.detach().float().permute(0, 2, 3, 1).contiguous()
    ops.nhwc_stem(x, w, scale, bias, y, k=3, stride=1, pad=1, act="leaky", resize_to=dst)
    got = ops.to_nchw(y)
    err = (got - want).abs()
    assert float(err.max()) < 4e-3 + 2e-3 * float(want.abs().max()), float(err.max())
    assert float(err.mean()) < 3e-4, float(err.mean())


def test_stem_weight_gradient_operand_is_the_resized_frame(native_lib):
    """dyk_frames_to_im2col32_resize against F.unfold of the interpolated frames (bf16 rounding of either side)."""
    import ctypes as C
    from dyk import _native as nat, ops
    g = torch.Generator().manual_seed(9)
    x = torch.randint(0, 256, (2, 3, 45, 61), dtype=torch.uint8, generator=g).to(DEV)
    H, W = 64, 80
    out = torch.empty((2, H, W, 32), dtype=torch.bfloat16, device=DEV)
    nat.call("dyk_frames_to_im2col32_resize", C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), 2, 45, 61, H, W,
             nat.DYK_BF16, 1, ops._stream())
    cols = F.unfold(_interp(x, (H, W)), 3, padding=1).view(2, 27, H, W).permute(0, 2, 3, 1)     # channel = ci*9 + r*3 + s
    assert torch.equal(out[..., 27:], torch.zeros_like(out[..., 27:]))
    err = (out[..., :27].float() - cols).abs()
    assert float(err.max()) < 2.0 ** -8, float(err.max())          # one bf16 ulp at values <= 1


def _model():
    import models
    from dyk import cfg_zoo
    torch.manual_seed(0)
    m = models.YOLO(cfg_zoo.materialize("kaist_yolov3.cfg"), (64, 96)).to(DEV)       # 3 -> 32 channel stem, single stream
    m.nc, m.gr = 1, 1.0
    m.hyp = {"box": 3.54, "cls": 37.4, "obj": 64.3, "cls_pw": 1.0, "obj_pw": 1.0, "iou_t": 0.20, "fl_gamma": 0.0}
    return m


def test_model_input_size_matches_pre_resized_frames(native_lib):
    """Eval forward: model(v, input_size=ns) on the original uint8 frames == the same model on F.interpolate(v / 255, ns)."""
    m = _model().eval()
    g = torch.Generator().manual_seed(1)
    v = torch.randint(0, 256, (2, 3, 80, 112), dtype=torch.uint8, generator=g).to(DEV)
    with torch.no_grad():
        io_f, _ = m(v, input_size=(64, 96))
        io_r, _ = m(_interp(v, (64, 96)))
    assert io_f.shape == io_r.shape == (2, 3 * (2 * 3 + 4 * 6 + 8 * 12), 6)
    rel = float((io_f - io_r).abs().max() / io_r.abs().max())
    assert rel < 1e-2, rel


def test_training_step_input_size_matches_pre_resized_frames(native_lib):
    """One multi-scale training step as in kaist_train_eval_utils.py:59-108 (resize, forward, loss, backward): fused resize
    vs pre-resized frames give the same loss and the same parameter gradients up to 16-bit rounding noise."""
    from build_utils.utils import compute_loss
    g = torch.Generator().manual_seed(2)
    v = torch.randint(0, 256, (2, 3, 96, 128), dtype=torch.uint8, generator=g).to(DEV)
    targets = torch.tensor([[0, 0, 0.5, 0.5, 0.3, 0.4], [1, 0, 0.3, 0.6, 0.2, 0.2]], device=DEV)
    res = []
    for fused in (True, False):
        m = _model().train()
        m.compute_dtype = torch.bfloat16
        pred = m(v, input_size=(64, 96)) if fused else m(_interp(v, (64, 96)))
        loss = compute_loss(pred, targets, m)
        loss = sum(loss.values()) if isinstance(loss, dict) else loss
        loss.backward()
        res.append((float(loss), [p.grad.clone() for p in m.parameters()]))
    (l1, g1), (l2, g2) = res
    assert abs(l1 - l2) < 2e-2 * abs(l2) + 1e-3, (l1, l2)
    stem = float((g1[0] - g2[0]).norm() / (g2[0].norm() + 1e-12))
    assert stem < 0.1, ("stem weight gradient", stem)
    tot = sum(float((a - b).norm()) for a, b in zip(g1, g2)) / sum(float(b.norm()) for b in g2)
    assert tot < 0.1, tot
