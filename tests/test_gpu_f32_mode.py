"""fp32-accurate mode (model.compute_dtype = torch.float32) on a B200 against the fp32 CPU oracle.

The north-star asks for box / conf / class within 1e-3 of the reference's fp32 path (reference models.py:279-315 run on
fp32 tensors).  Metrics, written out:
  * conf, class scores (sigmoid outputs in [0, 1]):          max |native - oracle|
  * boxes (xywh in pixels, up to ~640) and raw head logits:  max |native - oracle| / max(1, |oracle|)
Tolerance: TOL = 1e-3 — or, where two *exact fp32* implementations of the reference already differ by more than that,
2 x the difference between PyTorch's own CUDA fp32 path (cuDNN, TF32 off) and PyTorch's CPU fp32 path on the same
frames and weights.  That second clause is needed: on these seeded calibrated-random weights the 280-layer dyolov4 /
MobileNetV3 networks amplify fp32 accumulation-order noise to 2e-3 / 9e-3 in the scores (profiles/r02_drift_table.txt),
so no implementation can promise 1e-3 against a CPU run there; the Darknet53 models (configs[0], configs[1]) sit at 1e-4.
Convolutions run on the tcgen05 kernels as 3-way bf16 split products with chunked fp32 accumulation (csrc/f32_path.cu)."""
import pytest
import torch
import torch.nn.functional as F

from dyk import cfg_zoo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-3


def _frames(dual, B, H, W, seed=7):
    g = torch.Generator().manual_seed(seed)
    v = torch.rand((B, 3, H, W), generator=g)
    l = torch.rand((B, 3, H, W), generator=g) if dual else None
    return v, l


def _rel_err(got, want):
    got, want = got.float().cpu(), want.float()
    return float(((got - want).abs() / want.abs().clamp_min(1.0)).max())


@pytest.mark.parametrize("k,stride,cin,cout,act", [(3, 1, 64, 128, "mish"), (1, 1, 72, 40, "linear"), (3, 2, 32, 64, "leaky"),
                                                   (1, 2, 16, 64, "hard-swish"), (3, 1, 256, 512, "leaky")])
def test_split_bf16_conv_has_fp32_accuracy(native_lib, k, stride, cin, cout, act):
    """One convolution + folded BN + activation against the same op in float64."""
    import torch.nn as nn
    from dyk import ops
    from dyk.ops import View
    torch.manual_seed(k * 100 + cin)
    conv = nn.Conv2d(cin, cout, k, stride, k // 2, bias=False)
    bn = nn.BatchNorm2d(cout).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1); bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 2.0)
    x = torch.randn(2, cin, 40, 56)
    actf = {"mish": F.mish, "linear": lambda t: t, "leaky": lambda t: F.leaky_relu(t, 0.1), "hard-swish": F.hardswish}[act]
    want = actf(bn.double()(conv.double()(x.double()))).float()
    conv, bn = conv.float().to(DEV), bn.float().to(DEV)
    xv = View(x.permute(0, 2, 3, 1).contiguous().to(DEV), 0, cin)
    Ho, Wo = want.shape[2], want.shape[3]
    y = ops.new_view(2, Ho, Wo, cout, torch.float32, DEV)
    scale, bias = ops.fold_bn(conv, bn)
    split = torch.empty(xv.npix * 6 * cin, dtype=torch.bfloat16, device=DEV)
    ops.f32_conv(xv, ops.f32_pack_conv_weight(conv.weight), scale, bias, y, split, k=k, stride=stride, pad=k // 2, act=act, cout=cout)
    torch.cuda.synchronize()
    got = y.buf.permute(0, 3, 1, 2).cpu()
    err = float(((got - want).abs() / want.abs().clamp_min(1.0)).max())
    assert err < 4e-6, (k, stride, cin, cout, act, err)


@pytest.mark.parametrize("name,B,H,W", [("kaist_yolov3.cfg", 2, 512, 640), ("kaist_dyolov3_add_sl.cfg", 2, 512, 640),
                                        ("kaist_dyolov4_fshare_global_concat_se3.cfg", 2, 512, 640),
                                        ("kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg", 4, 512, 640),
                                        ("kaist_dyolov3_add_sl.cfg", 3, 96, 160)])
def test_model_fp32_mode_within_1e3_of_fp32_oracle(native_lib, name, B, H, W):
    import models
    from oracle import darknet_ref as dr
    from oracle import weights as ow
    path = cfg_zoo.materialize(name)
    ref = dr.DarknetRef(path)
    st = ow.make_calibrated_state(ref, seed=0)
    m = models.YOLO(path, (H, W))
    m.load_state_dict(st, strict=True)
    m = m.to(DEV).eval()
    m.compute_dtype = torch.float32
    dual = "second_index" in ref.net
    v, l = _frames(dual, B, H, W)
    def errors(io, p):
        io = io.float().cpu()
        return dict(box=_rel_err(io[..., :4], io_ref[..., :4]), score=float((io[..., 4:] - io_ref[..., 4:]).abs().max()),
                    logit=max(_rel_err(a, b) for a, b in zip(p, p_ref)))

    with torch.no_grad():
        io_ref, p_ref = ref.forward(st, v, l)
        io, p = m(v.to(DEV), l.to(DEV)) if dual else m(v.to(DEV))
        io2, _ = m(v.to(DEV), l.to(DEV)) if dual else m(v.to(DEV))           # second call = CUDA-graph replay
        # the reference's own fp32 arithmetic on this GPU (same torch ops on CUDA tensors, TF32 off): the noise floor
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        st_gpu = {k_: t.to(DEV) for k_, t in st.items()}
        io_g, p_g = ref.forward(st_gpu, v.to(DEV), l.to(DEV) if dual else None)
    torch.cuda.synchronize()
    assert torch.equal(io, io2)
    got, floor = errors(io, p), errors(io_g, p_g)
    for key in got:
        assert got[key] <= max(TOL, 2.0 * floor[key]), dict(model=name, metric=key, native=got, torch_cuda_fp32=floor)
    # uint8 frames (the /255 of the callers fused into the stem) are the same fp32 values
    v8 = (v * 255).round().to(torch.uint8)
    l8 = (l * 255).round().to(torch.uint8) if dual else None
    with torch.no_grad():
        io_u8, _ = m(v8.to(DEV), l8.to(DEV)) if dual else m(v8.to(DEV))
        io_f, _ = m((v8.float() / 255.0).to(DEV), (l8.float() / 255.0).to(DEV)) if dual else m((v8.float() / 255.0).to(DEV))
    assert torch.equal(io_u8, io_f)
