"""Whole-model training parity on a B200: one seeded optimisation step of models.YOLO (training mode: native forward with
batch-statistics BatchNorm + native backward inside one autograd node) against the oracle.

Three kinds of evidence, because gradients of a 16-bit run cannot be compared with an fp32 run element by element —
every leaky / relu kink that a 5e-4 rounding difference pushes across zero changes that element's derivative by 0.9
(measured: 3-5 % relative L2 error per parameter on a 10-layer net, decorrelation on the 100-layer random nets):
  (1) TIGHT, teacher-forced: every convolution block's dW / dgamma / dbeta / dx recomputed by the oracle from the
      *native* tensors of that block (oracle/layerwise.compare_backward) — must agree to 1e-3 (fp16) / 6e-3 (bf16);
  (2) end to end on a small net: loss, head logits, running statistics tight; parameter gradients within the kink noise;
  (3) structural: every parameter receives a finite gradient, two identical steps are bit-identical (deterministic
      reductions), the optimizer step changes what eval mode computes.
"""
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TINY = str(Path(__file__).resolve().parent / "data" / "tiny_yolov3_train.cfg")
TINY_MOBILE = str(Path(__file__).resolve().parent / "data" / "tiny_mobile_inc_train.cfg")


def _teacher_forced(tol):
    from oracle import layerwise
    from oracle.train_check import LAST
    rows = layerwise.compare_backward(LAST["plan"], LAST["frames"])
    assert len(rows) > 5
    for r in rows:
        for k, v in r.items():
            if k != "layer":
                assert v < tol, ("teacher-forced backward", r)
    return rows


def _teacher_forced_aux(tol, expect):
    """Fusion weights of the weighted shortcuts and SE fc parameters: native gradients against autograd of the same op on
    the native operands (oracle/layerwise.compare_backward_aux)."""
    from oracle import layerwise
    from oracle.train_check import LAST
    rows = layerwise.compare_backward_aux(LAST["plan"])
    kinds = sorted({r["kind"] for r in rows})
    assert kinds == sorted(expect), (kinds, expect)
    for r in rows:
        for k, v in r.items():
            if k not in ("layer", "kind"):
                assert v < tol, ("teacher-forced backward (shortcut / se parameters)", r)
    return rows


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_tiny_net_training_step(native_lib, dtype):
    from oracle.train_check import run
    rows, stats, fwd, loss, loss_o = run(TINY, 64, 96, 2, dtype)
    eps = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    assert abs(loss - loss_o) < 4 * eps * abs(loss_o) + 1e-3, (loss, loss_o)
    assert max(fwd) < 4 * eps, ("head logits", fwd)
    assert max(e for _, e in stats) < 2 * eps, ("BatchNorm running statistics", stats)
    assert all(r is not None for _, r, _ in rows), "a parameter received no gradient"
    worst = max(r for _, r, _ in rows)
    assert worst < (0.15 if dtype == torch.float16 else 0.6), ("end-to-end gradient error beyond the kink noise", worst)
    _teacher_forced(1e-3 if dtype == torch.float16 else 6e-3)


@pytest.mark.parametrize("name,dtype", [("kaist_dyolov3_add_sl.cfg", torch.float16),
                                        ("kaist_dyolov4_fshare_global_concat_se3.cfg", torch.bfloat16)])
def test_baseline_models_training_step(native_lib, name, dtype):
    """The two dense BASELINE models at 128x160, batch 2: all 393 / 568 parameter tensors get finite gradients, running
    statistics follow the oracle, and every convolution block passes the teacher-forced backward check."""
    from oracle.train_check import run
    rows, stats, fwd, loss, loss_o = run(name, 128, 160, 2, dtype)
    assert all(r is not None for _, r, _ in rows), "a parameter received no gradient"
    assert all(r == r and r < 10 for _, r, _ in rows), "non-finite gradient"
    assert max(e for _, e in stats) < (0.05 if dtype == torch.float16 else 0.3), ("running statistics drift", max(e for _, e in stats))
    _teacher_forced(1e-3 if dtype == torch.float16 else 6e-3)
    # the parameters outside the convolution blocks: 3 fusion weights (dyolov3_add_sl) / 4 fusion weights + 3 SE blocks (dyolov4)
    _teacher_forced_aux(2e-3 if dtype == torch.float16 else 1e-2, ["shortcut.w"] if "dyolov3" in name else ["se", "shortcut.w"])


def test_training_is_deterministic_and_updates_eval(native_lib):
    import models
    from dyk import cfg_zoo
    path = cfg_zoo.materialize("kaist_dyolov3_add_sl.cfg")
    g = torch.Generator().manual_seed(11)
    v = torch.rand((2, 3, 96, 128), generator=g).to(DEV)
    l = torch.rand((2, 3, 96, 128), generator=g).to(DEV)

    def one_step():
        torch.manual_seed(3)
        m = models.YOLO(path, (96, 128)).to(DEV).train()
        opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9)
        p = m(v, l)
        assert all(t.requires_grad for t in p) and [tuple(t.shape[1:]) for t in p] == [(3, 3, 4, 6), (3, 6, 8, 6), (3, 12, 16, 6)]
        loss = sum((t ** 2).mean() for t in p)
        loss.backward()
        grads = [q.grad.clone() for q in m.parameters()]
        opt.step()
        m.eval()
        with torch.no_grad():
            io, _ = m(v, l)
        return grads, io, m

    g1, io1, m1 = one_step()
    g2, io2, _ = one_step()
    assert all(torch.equal(a, b) for a, b in zip(g1, g2)), "backward is not bit-reproducible"
    assert torch.equal(io1, io2)
    assert all(torch.isfinite(a).all() for a in g1)
    # a second training step on the same model works (plan reuse) and changes the eval output (weights + running stats)
    m1.train()
    opt = torch.optim.SGD(m1.parameters(), lr=1e-2)
    sum((t ** 2).mean() for t in m1(v, l)).backward()
    opt.step()
    m1.eval()
    with torch.no_grad():
        io3, _ = m1(v, l)
    assert not torch.equal(io1, io3)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_tiny_mobile_inception_training_step(native_lib, dtype):
    """Depthwise 3x3 / 5x5 (stride 1 and 2), hard-swish / relu / relu6, SE, DepthwiseSeparableConv2d and Inception blocks."""
    from oracle.train_check import run
    rows, stats, fwd, loss, loss_o = run(TINY_MOBILE, 64, 96, 2, dtype)
    eps = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    assert max(fwd) < 8 * eps, ("head logits", fwd)
    assert max(e for _, e in stats) < 2 * eps, ("BatchNorm running statistics", stats)
    assert all(r is not None for _, r, _ in rows), "a parameter received no gradient"
    assert all(r == r and r < 10 for _, r, _ in rows), "non-finite gradient"
    _teacher_forced(1e-3 if dtype == torch.float16 else 6e-3)


def test_mobilenetv3_dual_training_step(native_lib):
    """BASELINE config 4's model (dual MobileNetV3 + FSNet + concat-SE + depthwise-separable PANet) at 128x160, batch 2:
    all 524 parameter tensors get finite gradients and every convolution block (dense and depthwise) passes the
    teacher-forced backward check."""
    from oracle.train_check import run
    rows, stats, fwd, loss, loss_o = run("kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg", 128, 160, 2, torch.bfloat16)
    assert all(r is not None for _, r, _ in rows), "a parameter received no gradient"
    assert all(r == r and r < 10 for _, r, _ in rows), "non-finite gradient"
    _teacher_forced(6e-3)


def test_training_limits_raise(native_lib):
    import models
    from dyk import _native
    cfg = Path(TINY).read_text().replace("filters=16\nsize=1", "filters=16\ngroups=2\nsize=1")
    assert "groups=2" in cfg
    path = Path("/tmp/dyk_tiny_grouped.cfg")
    path.write_text(cfg)
    m = models.YOLO(str(path), (64, 96)).to(DEV).train()
    x = torch.rand((1, 3, 64, 96), device=DEV)
    with pytest.raises(_native.NativeError):
        m(x)         # grouped (non-depthwise) convolutions: no kernels, and no silent fallback
