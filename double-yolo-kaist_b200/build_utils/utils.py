"""Hot-path part of the reference's build_utils/utils.py: ``non_max_suppression`` (utils.py:387-464),
``xywh2xyxy`` (:50-57), ``get_yolo_layers`` (:467-469), and the training loss ``compute_loss`` / ``build_targets``
(:209-384), which runs as fused native kernels (dyk/loss.py, csrc/yolo_loss.cu).

``non_max_suppression`` keeps the reference's signature and its list-of-(n,6)-tensors-or-None return
contract but processes the whole batch with three native kernel launches (csrc/nms.cu) instead of a
per-image Python loop around torchvision.ops.nms.  See INTEGRATION.md for how to bind these functions into an
unmodified reference checkout.

Deliberate difference: the reference aborts after a 10 s wall-clock budget and leaves later images as
None (utils.py:400,461-462), which makes its output timing dependent; the batched kernels have no such
guard (they finish a 128-image batch in well under a millisecond budget of that order).
"""
import ctypes as C

import torch

from dyk import _native as nat
from dyk import ops as _ops
from dyk.loss import build_targets, compute_loss

__all__ = ['non_max_suppression', 'xywh2xyxy', 'get_yolo_layers', 'compute_loss', 'build_targets', 'scale_coords',
           'clip_coords']

_workspaces = {}


def _boxes_inplace(coords, pad_x, pad_y, gain, img_shape):
    _ops._require_cuda(coords, "scale_coords / clip_coords")
    if coords.dtype != torch.float32 or coords.dim() != 2 or coords.shape[1] < 4 or (coords.shape[0] and coords.stride(1) != 1):
        raise ValueError("boxes must be a float32 (n, >= 4) tensor with unit column stride")
    with torch.cuda.device(coords.device):
        nat.call("dyk_scale_coords", C.c_void_p(coords.data_ptr()), coords.stride(0) if coords.shape[0] else 4, coords.shape[0],
                 float(pad_x), float(pad_y), float(gain), float(img_shape[1]), float(img_shape[0]),
                 torch.cuda.current_stream().cuda_stream)
    nat.count_launches()
    return coords


def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None):
    """Rescales xyxy boxes from the network input size (img1_shape) to the original image (img0_shape) and clips them, in
    place, like the reference (utils.py:60-84) — one native launch instead of four indexed tensor ops + four clamps."""
    if ratio_pad is None:
        gain = max(img1_shape) / max(img0_shape)
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    return _boxes_inplace(coords, pad[0], pad[1], gain, img0_shape)


def clip_coords(boxes, img_shape):
    """Clips xyxy boxes to (height, width) in place (utils.py:87-92)."""
    _boxes_inplace(boxes, 0.0, 0.0, 1.0, img_shape)


def get_yolo_layers(model):
    return [i for i, x in enumerate(model.module_defs) if x['type'] == 'yolo']


def xywh2xyxy(x):
    """(cx, cy, w, h) -> (x1, y1, x2, y2); host-side helper for callers (reference utils.py:50-57).  The
    NMS kernel performs the same conversion on device."""
    y = x.clone() if isinstance(x, torch.Tensor) else x.copy()
    y[:, 0] = x[:, 0] - x[:, 2] / 2
    y[:, 1] = x[:, 1] - x[:, 3] / 2
    y[:, 2] = x[:, 0] + x[:, 2] / 2
    y[:, 3] = x[:, 1] + x[:, 3] / 2
    return y


def nms_raw(prediction, conf_thres, iou_thres, multi_label, classes, agnostic, max_num):
    """Launches the batched NMS; returns device tensors (out [B, max_num, 6], counts [B] int32)."""
    _ops._require_cuda(prediction, "non_max_suppression")
    if prediction.dim() != 3 or prediction.shape[2] < 6:
        raise ValueError("prediction must be (batch, rows, 5 + nc)")
    pred = prediction.detach()
    if pred.dtype != torch.float32 or not pred.is_contiguous():
        pred = pred.float().contiguous()
    B, rows, no = pred.shape
    nc = no - 5
    mask = 0
    if classes:
        for c in classes:
            if not 0 <= int(c) < 64:
                raise ValueError("class filter supports ids 0..63")
            mask |= 1 << int(c)
    ml = int(bool(multi_label) and nc > 1)
    need = nat.load().dyk_nms_workspace_bytes(B, rows, nc, ml)
    key = (pred.device, torch.cuda.current_stream().cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=pred.device)
        _workspaces[key] = ws
    out = torch.empty((B, max_num, 6), dtype=torch.float32, device=pred.device)
    counts = torch.empty((B,), dtype=torch.int32, device=pred.device)
    nat.call("dyk_nms_batched", C.c_void_p(pred.data_ptr()), B, rows, nc, float(conf_thres), float(iou_thres), ml,
             C.c_uint64(mask), int(bool(agnostic)), int(max_num), C.c_void_p(out.data_ptr()),
             C.c_void_p(counts.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(),
             torch.cuda.current_stream().cuda_stream)
    nat.count_launches(3)
    return out, counts


def non_max_suppression(prediction, conf_thres=0.1, iou_thres=0.6, multi_label=True, classes=None, agnostic=False,
                        max_num=100):
    """Batched NMS on (batch, rows, 5 + nc) predictions.  Returns a list with one (n, 6) tensor
    (x1, y1, x2, y2, conf, cls) per image, or None where nothing survives (reference utils.py:387-464)."""
    if prediction.shape[0] == 0:
        return []
    out, counts = nms_raw(prediction, conf_thres, iou_thres, multi_label, classes, agnostic, max_num)
    n = counts.tolist()  # one device->host sync for the whole batch
    return [out[i, :k] if k > 0 else None for i, k in enumerate(n)]
