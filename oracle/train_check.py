"""ORACLE — TEST INFRASTRUCTURE ONLY.  One seeded training step of the native path next to the oracle's autograd.

Used by tests/test_gpu_train_model.py and tools/train_parity.py.  Loss = sum_i <p_i, R_i> with fixed random R_i (a
linear probe of every head output), so d loss / d p is known and identical on both sides.
"""
from __future__ import annotations

import os
from pathlib import Path

import torch

ROUND = os.environ.get("DYK_PARITY_FP32", "0") != "1"
LAST = {}


def run(cfg, H, W, B, dt, seed=5):
    """Returns (per-parameter rows (name, rel L2 error, |grad|), running-stat errors, head rel errors, loss native,
    loss oracle).  LAST["plan"], LAST["frames"], LAST["model"] keep the native objects for layerwise.compare_backward."""
    import models
    from dyk import cfg_zoo
    from oracle import darknet_ref as dr
    from oracle import weights as ow
    path = cfg if cfg.endswith(".cfg") and Path(cfg).exists() else cfg_zoo.materialize(cfg)
    ref = dr.DarknetRef(path)
    st = ow.fill_state(ref.shapes, seed=0)
    m = models.YOLO(path, (H, W))
    m.load_state_dict(st, strict=True)
    m = m.cuda().train()
    m.compute_dtype = dt
    dual = "second_index" in ref.net
    g = torch.Generator().manual_seed(seed)
    v = torch.rand((B, 3, H, W), generator=g)
    l = torch.rand((B, 3, H, W), generator=g) if dual else None
    p = m(v.cuda(), l.cuda()) if dual else m(v.cuda())
    R = [torch.randn(t.shape, generator=g) for t in p]
    loss = sum((a * r.cuda()).sum() for a, r in zip(p, R))
    loss.backward()
    torch.cuda.synchronize()
    LAST["plan"], LAST["frames"], LAST["model"] = m._train_plans.last_plan, (v, l), m
    # oracle: fp32 autograd on CPU, train-mode BN
    st_o = {k: (t.clone().requires_grad_(True) if t.dtype.is_floating_point and not k.endswith(("running_mean", "running_var"))
                else t.clone()) for k, t in st.items()}
    # storage-rounding model of the training plan: conv outputs (pre-BN) and every layer output rounded to the
    # 16-bit type, fp32 head logits; keeps the activation kinks (leaky / relu) on the same side as the native run
    heads = {i for i, d in enumerate(ref.defs) if d["type"] == "convolutional" and not ref.meta[i]["bn"]}
    p_o = ref.forward(st_o, v, l, training=True, round_dtype=dt if ROUND else None, unrounded=heads)
    loss_o = sum((a * r).sum() for a, r in zip(p_o, R))
    loss_o.backward()
    rows = []
    for name, prm in m.named_parameters():
        want = st_o[name].grad
        got = prm.grad.detach().float().cpu() if prm.grad is not None else None
        if want is None or got is None:
            rows.append((name, None, None))
            continue
        den = float(want.norm()) + 1e-20
        rows.append((name, float((got - want).norm()) / den, den))
    stats = []
    for name, buf in m.named_buffers():
        if name.endswith(("running_mean", "running_var")):
            want = st_o[name]
            stats.append((name, float((buf.detach().cpu() - want).norm()) / (float(want.norm()) + 1e-20)))
    fwd = [float((a.detach().cpu() - b.detach()).norm() / (b.detach().norm() + 1e-20)) for a, b in zip(p, p_o)]
    return rows, stats, fwd, float(loss.detach().cpu()), float(loss_o.detach())


