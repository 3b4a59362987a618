"""Native training loss behind the reference's `compute_loss(p, targets, model)` / `build_targets(p, targets, model)`
(reference build_utils/utils.py:209-384).  One autograd node: its forward runs target matching, box / objectness / class
losses AND their gradients on the device (csrc/yolo_loss.cu); its backward only scales the stored gradients by the
upstream gradients of the three loss outputs.  Nothing synchronises with the host: the number of matched targets lives
in device memory; bad targets (outside the batch / grid, class id >= nc) set a device flag that is read back
asynchronously and raised at the next call.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as nat
from .ops import _p, _require_cuda, _stream

_pending = []      # (event, pinned int tensor) of earlier calls, checked without blocking


def _check_pending(block=False):
    while _pending:
        ev, host = _pending[0]
        if not block and not ev.query():
            return
        ev.synchronize()
        _pending.pop(0)
        code = int(host[0])
        if code == 1:
            raise IndexError("compute_loss: a target indexes outside the batch or the grid (image index / xywh out of range)")
        if code == 2:
            raise AssertionError("compute_loss: a target class id is >= model.nc")


def _unwrap(model):
    return model.module if type(model) in (torch.nn.parallel.DataParallel, torch.nn.parallel.DistributedDataParallel) else model


def _head_config(model):
    m = _unwrap(model)
    anchors = [m.module_list[j].anchor_vec for j in m.yolo_layers]
    return m, anchors


def _prep_targets(targets, device):
    t = targets.detach().to(device=device, dtype=torch.float32).contiguous()
    if t.dim() != 2 or (t.shape[0] and t.shape[1] != 6):
        raise ValueError(f"targets must be (n, 6) = image, class, x, y, w, h; got {tuple(targets.shape)}")
    return t


def _match(p_i, t, anchor_vec, iou_t):
    dev = p_i.device
    nt, na = t.shape[0], anchor_vec.shape[0]
    B, na_p, ny, nx, no = p_i.shape
    if na_p != na:
        raise ValueError(f"head has {na_p} anchors but the yolo layer defines {na}")
    cap = max(na * nt, 1)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    idx = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    tbox = torch.empty((cap, 4), dtype=torch.float32, device=dev)
    tcls = torch.empty(cap, dtype=torch.int32, device=dev)
    av = anchor_vec.detach().to(device=dev, dtype=torch.float32).contiguous()
    nat.call("dyk_yolo_build_targets", _p(t) if nt else None, nt, _p(av), na, ny, nx, float(iou_t), _p(count), _p(idx), _p(tbox),
             _p(tcls), _stream())
    nat.count_launches()
    return count, idx, tbox, tcls, av


class _YoloLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, t, *p):
        dev = p[0].device
        acc = torch.zeros(3, dtype=torch.float32, device=dev)
        out = torch.empty(3, dtype=torch.float32, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        lib = nat.load()
        dps = []
        hyp = cfg["hyp"]
        for i, (pi, anchor_vec) in enumerate(zip(p, cfg["anchors"])):
            B, na, ny, nx, no = pi.shape
            count, idx, tbox, tcls, av = _match(pi, t, anchor_vec, hyp["iou_t"])
            max_pos = na * t.shape[0]
            cells = B * na * ny * nx
            ws = torch.empty(int(lib.dyk_yolo_loss_workspace_floats(max_pos, no, cells)), dtype=torch.float32, device=dev)
            dp = torch.empty_like(pi)
            nat.call("dyk_yolo_loss_head", _p(pi), _p(dp), B, na, ny, nx, no, _p(count), _p(idx), _p(tbox), _p(tcls), _p(av),
                     max_pos, int(cfg["v4"]), int("ciou" in hyp), float(cfg["gr"]), float(hyp["obj_pw"]), float(hyp["cls_pw"]),
                     float(hyp["box"]), float(hyp["obj"]), float(hyp["cls"]), _p(acc), int(i == len(p) - 1), _p(out), _p(err),
                     _p(ws), _stream())
            nat.count_launches(5)
            dps.append(dp)
        host = torch.empty(1, dtype=torch.int32, pin_memory=True)
        host.copy_(err, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        _pending.append((ev, host))
        ctx.dps = dps
        return out[0:1], out[1:2], out[2:3]

    @staticmethod
    def backward(ctx, g_box, g_obj, g_cls):
        dps = ctx.dps
        dev = dps[0].device
        up = torch.zeros(3, dtype=torch.float32, device=dev)
        for k, g in enumerate((g_box, g_obj, g_cls)):
            if g is not None:
                up[k:k + 1] = g.reshape(1)          # device-to-device copy of a scalar, no sync
        grads = []
        for dp in dps:
            o = torch.empty_like(dp)
            nat.call("dyk_yolo_loss_scale_grad", _p(dp), _p(o), dp.numel(), dp.shape[-1], _p(up), _stream())
            nat.count_launches()
            grads.append(o)
        return (None, None, *grads)


def compute_loss(p, targets, model):
    """Same contract as the reference's compute_loss: returns {"box_loss", "obj_loss", "class_loss"} (shape-(1,) tensors
    attached to the graph of the head tensors p)."""
    _check_pending()
    m, anchors = _head_config(model)
    if len(p) != len(anchors):
        raise ValueError(f"{len(p)} head tensors for {len(anchors)} yolo layers")
    _require_cuda(p[0], "compute_loss")
    hyp = m.hyp
    if float(hyp.get("fl_gamma", 0.0)) > 0:
        raise nat.NativeError("compute_loss: focal loss (fl_gamma > 0) has no native kernel")
    nc = int(m.nc)
    ps = []
    for pi in p:
        if pi.dim() != 5 or pi.shape[-1] != 5 + nc:
            raise ValueError(f"head tensor {tuple(pi.shape)} does not match model.nc = {nc}")
        ps.append(pi.float().contiguous())
    t = _prep_targets(targets, ps[0].device)
    cfg = dict(hyp=hyp, anchors=anchors, gr=float(m.gr), v4="yolov4" in m.cfg)
    with torch.cuda.device(ps[0].device):
        lbox, lobj, lcls = _YoloLoss.apply(cfg, t, *ps)
    return {"box_loss": lbox, "obj_loss": lobj, "class_loss": lcls}


def build_targets(p, targets, model):
    """Same contract as the reference's build_targets: (tcls, tbox, indices, anch) per head, int64 indices.
    (Reads the positive count back, so this diagnostic entry point synchronises; compute_loss does not.)"""
    m, anchors = _head_config(model)
    _require_cuda(p[0], "build_targets")
    t = _prep_targets(targets, p[0].device)
    tcls, tbox, indices, anch = [], [], [], []
    with torch.cuda.device(p[0].device):
        for pi, anchor_vec in zip(p, anchors):
            count, idx, tb, tc, av = _match(pi, t, anchor_vec, m.hyp["iou_t"])
            n = int(count.item())
            idx = idx[:n].long()
            indices.append((idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]))
            tbox.append(tb[:n])
            tcls.append(tc[:n].long())
            anch.append(av[idx[:, 1]])
    return tcls, tbox, indices, anch
