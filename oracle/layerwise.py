"""ORACLE — TEST INFRASTRUCTURE ONLY.  Teacher-forced, layer-by-layer comparison of a native plan with the oracle.

Used by tests/test_gpu_model.py, __graft_entry__.smoke() and tools/layer_parity.py.  The native model must have been
built with DYK_NO_REUSE=1 (no buffer recycling) and run once, so that every materialised layer output is still in
its buffer.  The oracle (oracle/darknet_ref.py, run as the storage-rounding model of the plan) then recomputes every
layer from the *native* inputs of that layer; see DarknetRef.forward(teacher=...) for why only this comparison can
be tight.
"""
from __future__ import annotations

import torch

ULP = {torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}


def native_layers(model):
    """({layer index: NCHW fp32 CPU tensor} of every materialised layer output, {layers never stored in 16 bits})."""
    from dyk import ops
    from dyk.ops import View
    plan = model._plans.last_plan
    got, unrounded = {}, set()
    for i, v in enumerate(plan.layer_vals):
        if v is None or v.ext:
            continue
        if (v.view is None and not getattr(v, "virtual", False)) or v.f32:
            unrounded.add(i)      # (a `virtual` value is formed inside its consumer but IS rounded to the storage type there)
        if v.view is None:
            continue
        vw = v.view
        if v.f32:
            got[i] = vw.buf[..., vw.c_off:vw.c_off + vw.C].permute(0, 3, 1, 2).float().cpu().contiguous()
        else:
            got[i] = ops.to_nchw(View(vw.buf, vw.c_off, vw.C)).cpu()
    return got, unrounded


def gated_convs(model):
    """{conv layer index: se layer index} of the 1x1 convolutions that fold a SqueezeExcitation gate into per-image weights."""
    plan = model._plans.last_plan
    return {op.layer: op.gate_of.layer for op in plan.ops if getattr(op, "kind", "") == "conv" and getattr(op, "gate_of", None) is not None}


def compare(model, ref, st, v, l, dtype):
    """Returns a list of per-layer records {layer, type, max_diff, over_2ulp, n, frac_diff, rel_rms}.
    v, l: the fp32 NCHW frames (already /255) the native run was given (l None for single-stream cfgs)."""
    got, unrounded = native_layers(model)
    with torch.no_grad():
        plan = model._plans.last_plan
        gates = {op.layer: op.gate.float().cpu() for op in plan.ops if getattr(op, "kind", "") == "se" and getattr(op, "gate_only", False)}
        _, every = ref.forward(st, v, l, keep_layers=True, round_dtype=dtype, unrounded=unrounded, teacher=got,
                               gated=gated_convs(model), teacher_gates=gates)
    ulp = ULP[dtype]
    rows = []
    for i, nat_out in sorted(got.items()):
        want = every[i]
        if want.shape != nat_out.shape:
            rows.append(dict(layer=i, type=ref.defs[i]["type"], shape_mismatch=(tuple(want.shape), tuple(nat_out.shape))))
            continue
        rms = float(want.pow(2).mean().sqrt())
        floor = torch.full_like(want, 0.25 * rms)
        if ref.defs[i]["type"] == "convolutional" and i in unrounded:     # fp32 head logits: accumulation order only
            tol = 2e-3 * torch.maximum(want.abs(), torch.full_like(want, rms))
        else:
            tol = 2.0 * ulp * torch.maximum(want.abs(), floor) + 1e-6
        diff = (nat_out - want).abs()
        rows.append(dict(layer=i, type=ref.defs[i]["type"], max_diff=float(diff.max()), over_2ulp=int((diff > tol).sum()),
                         n=diff.numel(), frac_diff=float((diff > 0.25 * ulp * torch.maximum(want.abs(), floor)).float().mean()),
                         rel_rms=float(diff.pow(2).mean().sqrt()) / (rms + 1e-12)))
    return rows


def assert_layerwise(rows, n_layers, what=""):
    checked = 0
    for r in rows:
        assert "shape_mismatch" not in r, (what, r)
        assert r["over_2ulp"] == 0, (what, "layer differs from the oracle by more than 2 ulp of the storage type", r)
        assert r["frac_diff"] < 0.05, (what, "too many elements differ", r)
        checked += 1
    assert checked > n_layers // 3, (what, "too few layers were compared", checked)


def _nchw(view):
    return view.buf[..., view.c_off:view.c_off + view.C].permute(0, 3, 1, 2).float().cpu().contiguous()


def compare_backward(plan, frames):
    """Teacher-forced check of the training backward, convolution block by convolution block.

    After one native forward + backward (plan = model._train_plans.last_plan) every block's *native* input x, the native
    gradient dy of its output and the native parameter gradients are still in the plan's buffers.  For each
    [convolutional] block the oracle recomputes y = act(BN_train(conv(x, W))) in fp32 torch from the native x (weights
    rounded to the 16-bit type like the tensor-core kernel sees them), back-propagates the native dy through it with
    autograd and compares dW, dgamma, dbeta (and dx where this block is the only writer of the input gradient).
    Evaluating every block on the native tensors keeps the activation kinks on the same side, which an end-to-end
    comparison of a 16-bit run with an fp32 run cannot (see tools/train_parity.py).
    frames: (visible, lwir) NCHW fp32 CPU tensors as given to the model (for the stem blocks).
    Returns a list of dict(layer, dW, dgamma, dbeta, dx) with relative L2 errors (None where not applicable)."""
    import torch.nn.functional as F
    from oracle.darknet_ref import activation
    flat = plan.last_flat
    rows = []
    for st in plan.convs:
        op, conv, bn = st["op"], st["conv"], st["bn"]
        k, s, p = st["k"], st["s"], st["p"]
        if st["stem"]:
            x = (frames[0] if op.src is plan.img0 else frames[1]).clone()
            w = conv.weight.detach().float().cpu().clone()
        elif conv.groups > 1:           # depthwise: the native kernels read the fp32 filter
            x = _nchw(op.src.view)
            w = conv.weight.detach().float().cpu().clone()
        else:
            x = _nchw(op.src.view)
            w = conv.weight.detach().to(plan.dtype).float().cpu().clone()
        dy_full = _nchw(st["dy_view"])
        if bn is not None:
            # BN + activation on the *native* pre-BN tensor z (so every activation kink sits where the native run had it),
            # then the convolution gradients from the 16-bit-rounded dz, as the native wgrad / dgrad kernels see it
            g = bn.weight.detach().float().cpu().clone().requires_grad_(True)
            b = bn.bias.detach().float().cpu().clone().requires_grad_(True)
            z = _nchw(st["z"]).requires_grad_(True)
            zb = F.batch_norm(z, None, None, g, b, True, 0.0, bn.eps)
            if op.act == "mish":
                # storage-rounding model of the native BN backward for Mish: g = dy * mish'(zhat) is written to the 16-bit
                # dz buffer between its two passes (the reference under autocast rounds the same tensor between its
                # mish_backward and batch_norm_backward kernels)
                zb_d = zb.detach().requires_grad_(True)
                activation(zb_d, op.act).backward(dy_full)
                zb.backward(zb_d.grad, retain_graph=True)             # dgamma / dbeta: the per-channel sums use the fp32 g
                gg, bg = g.grad.clone(), b.grad.clone()
                z.grad = None
                zb.backward(zb_d.grad.to(plan.dtype).float())         # dz: from the stored (rounded) g
                g.grad, b.grad = gg, bg
            else:
                activation(zb, op.act).backward(dy_full)
            dz = z.grad.to(plan.dtype).float()
            bias = None
        else:
            g = b = None
            bias = conv.bias.detach().float().cpu().clone().requires_grad_(True)
            dz = dy_full[:, :conv.out_channels]
        w_grad = torch.nn.grad.conv2d_weight(x, w.shape, dz, stride=s, padding=p, groups=conv.groups)
        x_grad = None if st["stem"] else torch.nn.grad.conv2d_input(x.shape, w, dz, stride=s, padding=p, groups=conv.groups)
        bias_grad = None if bias is None else dz.sum((0, 2, 3))

        def rel(got, want):
            return float((got.cpu() - want).norm() / (want.norm() + 1e-20))

        row = dict(layer=op.layer, dW=rel(plan._pgrad(flat, conv.weight), w_grad))
        if bn is not None:
            row["dgamma"] = rel(plan._pgrad(flat, bn.weight), g.grad)
            row["dbeta"] = rel(plan._pgrad(flat, bn.bias), b.grad)
        else:
            row["dbias"] = rel(plan._pgrad(flat, conv.bias), bias_grad)
        if not st["stem"] and len(plan.consumers.get(id(op.src), [])) == 1 and st["gin"] is not None and not st["gin"][1]:
            row["dx"] = rel(_nchw(st["gin"][0]), x_grad)
        rows.append(row)
    return rows


def compare_backward_aux(plan):
    """Teacher-forced check of the parameter gradients that do not belong to a convolution block: the fusion weights `w` of
    weighted [shortcut]s (build_utils/layers.py:63-85) and the fc1 / fc2 parameters of [se] blocks (layers.py:175-190).
    As in compare_backward, the oracle recomputes the op in fp32 torch from the *native* operands, back-propagates the
    native gradient dy of the op's output with autograd and compares the parameter gradients the native backward wrote into
    the flat gradient buffer.  Returns [{layer, kind, <param>: relative L2 error}]."""
    import torch.nn.functional as F
    flat = plan.last_flat
    rows = []

    def rel(got, want):
        return float((got.detach().float().cpu().reshape(want.shape) - want).norm() / (want.norm() + 1e-20))

    for kind, op, dy_view in plan.aux_bwd:
        dy = _nchw(dy_view)
        if kind == "add":
            m = op.module
            if not m.weight:
                continue
            x, a = _nchw(op.x.view), _nchw(op.others[0].view)
            w = m.w.detach().float().cpu().clone().requires_grad_(True)
            ws = torch.sigmoid(w) * (2 / 2)
            (x * ws[0] + a * ws[1]).backward(dy)
            rows.append(dict(layer=op.layer, kind="shortcut.w", w=rel(plan._pgrad(flat, m.w), w.grad)))
        else:
            m = op.module
            x = _nchw(op.src.view)
            prm = {n: getattr(getattr(m, f), k).detach().float().cpu().clone().requires_grad_(True)
                   for n, (f, k) in dict(w1=("fc1", "weight"), b1=("fc1", "bias"), w2=("fc2", "weight"), b2=("fc2", "bias")).items()}
            s_ = F.adaptive_avg_pool2d(x, (1, 1))
            s_ = F.relu(F.conv2d(s_, prm["w1"], prm["b1"]))
            s_ = F.hardsigmoid(F.conv2d(s_, prm["w2"], prm["b2"]))
            (s_ * x).backward(dy)
            rows.append(dict(layer=op.layer, kind="se", fc1_w=rel(plan._pgrad(flat, m.fc1.weight), prm["w1"].grad),
                             fc1_b=rel(plan._pgrad(flat, m.fc1.bias), prm["b1"].grad),
                             fc2_w=rel(plan._pgrad(flat, m.fc2.weight), prm["w2"].grad),
                             fc2_b=rel(plan._pgrad(flat, m.fc2.bias), prm["b2"].grad)))
    return rows
