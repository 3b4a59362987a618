"""Fused multi-tensor optimizers for the training step.

Drop-in replacements for the optimizers the reference builds (train.py:85-91)::

    optimizer = dyk.optim.FusedSGD(pg, lr=hyp["lr0"], momentum=hyp["momentum"], weight_decay=hyp["weight_decay"], nesterov=True)
    optimizer = dyk.optim.FusedAdam(pg, lr=hyp["lr0"], betas=(hyp["momentum"], 0.999), weight_decay=hyp["weight_decay"])

and stepped the way it steps them — ``scaler.step(optimizer); scaler.update(); optimizer.zero_grad()``
(train_utils/kaist_train_eval_utils.py:103-108): both classes are ``torch.optim.Optimizer`` subclasses (param groups,
``state_dict``, LR schedulers work unchanged) and declare ``_step_supports_amp_scaling``, so ``GradScaler.step`` hands them
its device-side ``grad_scale`` / ``found_inf`` tensors and the un-scaling and the skip-on-overflow happen inside the kernel,
without a host synchronisation.

One native launch per param group updates every tensor of the group (csrc/optim.cu: a device table of
(param, grad, state) pointers, a block finds its tensor by binary search).  The table is rebuilt only when a pointer
changes; with ``dyk.train_plan`` the gradients of a step are views of one flat buffer, so usually it never does.
There is no eager fallback: CPU parameters raise.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as nat
from . import ops


class _FusedBase(torch.optim.Optimizer):
    _step_supports_amp_scaling = True      # GradScaler.step sets self.grad_scale / self.found_inf and calls step()
    _n_states = 1

    def _table(self, gi, params, grads, states):
        """Device descriptor table of param group `gi`, cached on the pointer signature (kept on the optimizer object, not
        in the param group, so that state_dict() stays what torch.optim would write)."""
        sig = tuple(t.data_ptr() for t in params) + tuple(t.data_ptr() for t in grads)
        cache = self.__dict__.setdefault("_dyk_tables", {}).setdefault(gi, {})
        if cache.get("sig") == sig:
            return cache["desc"], cache["blocks"]
        per = nat.load().dyk_optim_block_elems()
        rows, blocks = [], 0
        for i, (p, g) in enumerate(zip(params, grads)):
            st = states[i]
            rows.append([p.data_ptr(), g.data_ptr(), st[0].data_ptr(), st[1].data_ptr() if len(st) > 1 else 0, p.numel(), blocks])
            blocks += (p.numel() + per - 1) // per
        host = torch.tensor(rows, dtype=torch.int64).pin_memory()
        desc = host.to(params[0].device, non_blocking=True)
        cache.update(sig=sig, desc=desc, blocks=blocks, host=host)     # `host` kept alive until the async copy has run
        return desc, blocks

    def _collect(self, group):
        params, grads = [], []
        for p in group["params"]:
            if p.grad is None:
                continue
            if not p.is_cuda:
                raise nat.NativeError("dyk.optim: parameters must live on a B200 (no CPU fallback)")
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                raise nat.NativeError("dyk.optim: parameters and gradients must be dense float32 tensors")
            if not p.is_contiguous():
                raise nat.NativeError("dyk.optim: parameters must be contiguous")
            params.append(p)
            grads.append(p.grad if p.grad.is_contiguous() else p.grad.contiguous())
        return params, grads

    def _amp(self):
        gs, fi = getattr(self, "grad_scale", None), getattr(self, "found_inf", None)
        if gs is not None and not isinstance(gs, torch.Tensor):
            gs = None if gs == 1 else torch.tensor(float(gs), device=self.param_groups[0]["params"][0].device)
        if gs is not None:
            gs = gs.float().reshape(-1)[:1].contiguous()
        if fi is not None:
            fi = fi.float().reshape(-1)[:1].contiguous()
        return gs, fi


class FusedSGD(_FusedBase):
    """torch.optim.SGD(params, lr, momentum, dampening, weight_decay, nesterov) as one native launch per param group."""

    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        if lr < 0 or momentum < 0 or weight_decay < 0:
            raise ValueError("invalid hyper-parameter")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        gs, fi = self._amp()
        started = self.__dict__.setdefault("_dyk_started", set())
        for gi, group in enumerate(self.param_groups):
            params, grads = self._collect(group)
            if not params:
                continue
            states = []
            for p in params:
                st = self.state[p]
                if st.get("momentum_buffer") is None:
                    st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                states.append((st["momentum_buffer"],))
            # torch's first step sets buf = g; with zero-initialised buffers momentum*0 + (1-dampening)*g is the same
            # unless dampening != 0, which then needs the explicit first-step flag (and cannot be combined with a
            # GradScaler, whose skip decision only exists on the device)
            first = gi not in started and group["dampening"] != 0
            if first and fi is not None:
                raise nat.NativeError("dyk.optim.FusedSGD: dampening != 0 together with GradScaler is not supported")
            with torch.cuda.device(params[0].device):
                desc, blocks = self._table(gi, params, grads, states)
                nat.call("dyk_optim_sgd_multi", ops._p(desc), len(params), blocks, float(group["lr"]), float(group["momentum"]),
                         float(group["dampening"]), float(group["weight_decay"]), int(bool(group["nesterov"])), int(first),
                         ops._p(gs), ops._p(fi), ops._stream())
            nat.count_launches()
            started.add(gi)
        return loss


class FusedAdam(_FusedBase):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) (L2 weight decay, no amsgrad) as one native launch per group."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        gs, fi = self._amp()
        if fi is not None:
            # the bias corrections depend on the host-side step count; a skipped step would have to not count.  Read the
            # flag (one 4-byte D2H) only in this AMP + Adam combination
            if float(fi.item()) != 0.0:
                return loss
            fi = None
        for gi, group in enumerate(self.param_groups):
            params, grads = self._collect(group)
            if not params:
                continue
            states = []
            for p in params:
                st = self.state[p]
                if "exp_avg" not in st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                states.append((st["exp_avg"], st["exp_avg_sq"]))
            steps = {self.state[p]["step"] for p in params}
            if len(steps) != 1:
                raise nat.NativeError("dyk.optim.FusedAdam: parameters of one group must share their step count")
            b1, b2 = group["betas"]
            with torch.cuda.device(params[0].device):
                desc, blocks = self._table(gi, params, grads, states)
                nat.call("dyk_optim_adam_multi", ops._p(desc), len(params), blocks, float(group["lr"]), float(b1), float(b2),
                         float(group["eps"]), float(group["weight_decay"]), int(steps.pop()), ops._p(gs), ops._p(fi), ops._stream())
            nat.count_launches()
        return loss
