"""ORACLE — TEST INFRASTRUCTURE ONLY.  non_max_suppression restated in numpy fp32.

Follows the reference's build_utils/utils.py:387-464 step by step (xywh2xyxy :50-57), with the
torchvision.ops.nms call at :448 replaced by its published algorithm (torchvision 0.26
csrc/ops/cpu/nms_kernel.cpp: stable descending sort of the scores, greedy suppression of every later box
whose IoU with a kept box is strictly greater than the threshold; IoU = inter / (area_i + area_j - inter),
areas (x2-x1)*(y2-y1), all in the input dtype).  Pinned against the real reference + torchvision by
tests/golden/nms_*.npz (tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def xywh2xyxy(x: np.ndarray) -> np.ndarray:
    y = np.zeros_like(x)
    y[:, 0] = x[:, 0] - x[:, 2] / F32(2)
    y[:, 1] = x[:, 1] - x[:, 3] / F32(2)
    y[:, 2] = x[:, 0] + x[:, 2] / F32(2)
    y[:, 3] = x[:, 1] + x[:, 3] / F32(2)
    return y


def greedy_nms(boxes: np.ndarray, scores: np.ndarray, iou_thres: float, limit: int = None) -> np.ndarray:
    """Indices kept by torchvision.ops.nms(boxes, scores, iou_thres) (optionally only the first `limit`)."""
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-scores, kind="stable")
    b = boxes[order].astype(F32)
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    thr = float(iou_thres)  # torchvision's CPU kernel compares the fp32 IoU against the *double* threshold
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(order[i])
        if limit is not None and len(keep) >= limit:
            break
        if i + 1 == n:
            break
        xx1 = np.maximum(x1[i], x1[i + 1:])
        yy1 = np.maximum(y1[i], y1[i + 1:])
        xx2 = np.minimum(x2[i], x2[i + 1:])
        yy2 = np.minimum(y2[i], y2[i + 1:])
        w = np.maximum(F32(0), xx2 - xx1)
        h = np.maximum(F32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[i + 1:] - inter)
        suppressed[i + 1:] |= ovr.astype(np.float64) > thr
    return np.asarray(keep, dtype=np.int64)


def non_max_suppression(prediction: np.ndarray, conf_thres=0.1, iou_thres=0.6, multi_label=True, classes=None,
                        agnostic=False, max_num=100):
    """prediction (B, rows, 5+nc) fp32 -> list of (n, 6) fp32 arrays or None (utils.py:387-464)."""
    prediction = np.asarray(prediction, dtype=F32)
    min_wh, max_wh = 2, 4096
    nc = prediction.shape[2] - 5
    multi_label = bool(multi_label) and nc > 1
    output = [None] * prediction.shape[0]
    conf_t = F32(conf_thres)
    for xi in range(prediction.shape[0]):
        x = prediction[xi]
        x = x[x[:, 4] > conf_t]
        x = x[((x[:, 2:4] > min_wh) & (x[:, 2:4] < max_wh)).all(1)]
        if not x.shape[0]:
            continue
        x = x.copy()
        x[:, 5:] *= x[:, 4:5]
        box = xywh2xyxy(x[:, :4])
        if multi_label:
            i, j = np.nonzero(x[:, 5:] > conf_t)
            x = np.concatenate((box[i], x[i, j + 5][:, None], j.astype(F32)[:, None]), 1)
        else:
            j = x[:, 5:].argmax(1)
            conf = x[np.arange(x.shape[0]), j + 5]
            x = np.concatenate((box, conf[:, None], j.astype(F32)[:, None]), 1)
            sel = conf > conf_t
            x, j = x[sel], j[sel]
        if classes:
            x = x[np.isin(j, np.asarray(classes))]
        if not x.shape[0]:
            continue
        c = x[:, 5] * F32(0) if agnostic else x[:, 5]
        boxes = x[:, :4] + c[:, None] * F32(max_wh)
        keep = greedy_nms(boxes, x[:, 4], iou_thres, limit=max_num)
        output[xi] = x[keep[:max_num]]
    return output
