"""ORACLE — TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product path).

CPU restatement, in plain torch fp32 with autograd, of the reference's training loss for one batch:
  * build_utils/utils.py:305-384  build_targets  (anchor matching by wh-IoU > iou_t, anchor-major order of positives)
  * build_utils/utils.py:209-302  compute_loss   (GIoU / CIoU box loss, IoU-aware objectness BCE, class BCE when nc > 1)
  * build_utils/utils.py:95-138   bbox_iou       (xywh form, GIoU and CIoU branches)
  * build_utils/utils.py:166-172  wh_iou
Pinned against the real reference by tests/golden/loss_cases.npz (tests/golden/make_loss_golden.py runs the reference's
own functions on the same seeded inputs).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def wh_iou(wh1, wh2):
    a, b = wh1[:, None], wh2[None]
    inter = torch.min(a, b).prod(2)
    return inter / (a.prod(2) + b.prod(2) - inter)


def build_targets(p, targets, anchor_vecs, iou_t):
    """p: list of (B, na, ny, nx, no); targets (nt, 6) = image, class, x, y, w, h (normalised).
    Returns per head (tcls, tbox, (b, a, gj, gi), anchors-of-positives)."""
    out = []
    nt = targets.shape[0]
    for pi, anchors in zip(p, anchor_vecs):
        ny, nx = pi.shape[2], pi.shape[3]
        gain = torch.tensor([1.0, 1.0, nx, ny, nx, ny])
        t = targets * gain
        na = anchors.shape[0]
        if nt:
            keep = wh_iou(anchors, t[:, 4:6]) > iou_t                    # (na, nt)
            a = torch.arange(na).view(na, 1).repeat(1, nt)[keep]         # anchor-major order
            t = t.repeat(na, 1, 1)[keep]
        else:
            a = torch.zeros(0, dtype=torch.long)
        b, c = t[:, 0].long(), t[:, 1].long()
        gxy, gwh = t[:, 2:4], t[:, 4:6]
        gij = gxy.long()                                                 # truncation towards zero, like .long()
        out.append((c, torch.cat((gxy - gij, gwh), 1), (b, a, gij[:, 1], gij[:, 0]), anchors[a]))
    return out


def bbox_iou_xywh(box1, box2, ciou):
    """box1 (4, n) predicted, box2 (n, 4) target, both xywh; GIoU when not ciou (utils.py:95-138)."""
    box2 = box2.t()
    b1x1, b1x2 = box1[0] - box1[2] / 2, box1[0] + box1[2] / 2
    b1y1, b1y2 = box1[1] - box1[3] / 2, box1[1] + box1[3] / 2
    b2x1, b2x2 = box2[0] - box2[2] / 2, box2[0] + box2[2] / 2
    b2y1, b2y2 = box2[1] - box2[3] / 2, box2[1] + box2[3] / 2
    inter = (torch.min(b1x2, b2x2) - torch.max(b1x1, b2x1)).clamp(0) * (torch.min(b1y2, b2y2) - torch.max(b1y1, b2y1)).clamp(0)
    w1, h1 = b1x2 - b1x1, b1y2 - b1y1
    w2, h2 = b2x2 - b2x1, b2y2 - b2y1
    union = (w1 * h1 + 1e-16) + w2 * h2 - inter
    iou = inter / union
    cw = torch.max(b1x2, b2x2) - torch.min(b1x1, b2x1)
    ch = torch.max(b1y2, b2y2) - torch.min(b1y1, b2y1)
    if not ciou:
        c_area = cw * ch + 1e-16
        return iou - (c_area - union) / c_area
    c2 = cw ** 2 + ch ** 2 + 1e-16
    rho2 = ((b2x1 + b2x2) - (b1x1 + b1x2)) ** 2 / 4 + ((b2y1 + b2y2) - (b1y1 + b1y2)) ** 2 / 4
    v = (4 / math.pi ** 2) * torch.pow(torch.atan(w2 / h2) - torch.atan(w1 / h1), 2)
    with torch.no_grad():
        alpha = v / (1 - iou + v)
    return iou - (rho2 / c2 + v * alpha)


def compute_loss(p, targets, anchor_vecs, hyp, gr, nc, v4):
    """Returns (box_loss, obj_loss, class_loss) as shape-(1,) tensors, weighted by hyp['box'|'obj'|'cls'] (utils.py:209-302).
    hyp needs box, obj, cls, cls_pw, obj_pw, iou_t, fl_gamma (must be 0) and optionally the key 'ciou'."""
    assert hyp.get("fl_gamma", 0.0) == 0.0, "focal loss is not restated"
    lcls, lbox, lobj = torch.zeros(1), torch.zeros(1), torch.zeros(1)
    built = build_targets(p, targets, anchor_vecs, hyp["iou_t"])
    cls_pw, obj_pw = torch.tensor([hyp["cls_pw"]]), torch.tensor([hyp["obj_pw"]])
    for pi, (tcls, tbox, (b, a, gj, gi), anch) in zip(p, built):
        tobj = torch.zeros_like(pi[..., 0])
        nb = b.shape[0]
        if nb:
            ps = pi[b, a, gj, gi]
            if v4:
                pxy = ps[:, :2].sigmoid() * 2.0 - 0.5
                pwh = (ps[:, 2:4].sigmoid() * 2) ** 2 * anch
            else:
                pxy = ps[:, :2].sigmoid()
                pwh = ps[:, 2:4].exp().clamp(max=1e3) * anch
            pbox = torch.cat((pxy, pwh), 1)
            iou = bbox_iou_xywh(pbox.t(), tbox, "ciou" in hyp)
            lbox = lbox + (1.0 - iou).mean()
            tobj[b, a, gj, gi] = (1.0 - gr) + gr * iou.detach().clamp(0).type(tobj.dtype)    # duplicates: last write wins
            if nc > 1:
                t = torch.zeros_like(ps[:, 5:])
                t[range(nb), tcls] = 1.0
                lcls = lcls + F.binary_cross_entropy_with_logits(ps[:, 5:], t, pos_weight=cls_pw)
        lobj = lobj + F.binary_cross_entropy_with_logits(pi[..., 4], tobj, pos_weight=obj_pw)
    return lbox * hyp["box"], lobj * hyp["obj"], lcls * hyp["cls"]
