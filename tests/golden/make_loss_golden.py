"""Generates tests/golden/loss_cases.npz from the REAL reference (run in the build container, needs /root/reference):
the reference's own `compute_loss` / `build_targets` (build_utils/utils.py:209-384) on seeded head tensors and targets,
with the gradients of box + obj + class loss with respect to the head tensors from its autograd graph.

    python tests/golden/make_loss_golden.py
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import import_reference  # noqa: E402

ANCHORS = torch.tensor([[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]],
                       dtype=torch.float32)
CASES = [  # name, v4, ciou, nc, gr, obj_pw, cls_pw, nt, seed
    ("v3_giou", False, False, 1, 1.0, 1.0, 1.0, 9, 1),
    ("v4_ciou", True, True, 1, 0.5, 1.5, 1.0, 12, 2),
    ("v3_ciou_nc3", False, True, 3, 1.0, 1.0, 2.0, 10, 3),
    ("v4_giou_empty", True, False, 1, 1.0, 1.0, 1.0, 0, 4),
    ("v3_ciou_edges", False, True, 1, 0.3, 1.0, 1.0, 8, 5),       # labels on cell boundaries / image borders / unmatched sizes
    ("v4_giou_nc2_pw", True, False, 2, 0.7, 0.7, 1.3, 11, 6),
]
H, W, B = 128, 192, 2


def case_inputs(v4, nc, nt, seed):
    g = torch.Generator().manual_seed(seed)
    strides = [8, 16, 32] if v4 else [32, 16, 8]
    masks = [[0, 1, 2], [3, 4, 5], [6, 7, 8]] if v4 else [[6, 7, 8], [3, 4, 5], [0, 1, 2]]
    p = [torch.randn((B, 3, H // s, W // s, 5 + nc), generator=g) for s in strides]
    anchor_vecs = [ANCHORS[m] / s for m, s in zip(masks, strides)]
    t = torch.zeros((nt, 6))
    if nt:
        t[:, 0] = torch.randint(0, B, (nt,), generator=g).float()
        t[:, 1] = torch.randint(0, nc, (nt,), generator=g).float()
        t[:, 2:4] = torch.rand((nt, 2), generator=g) * 0.8 + 0.1
        t[:, 4] = torch.rand((nt,), generator=g) * 0.25 + 0.03
        t[:, 5] = torch.rand((nt,), generator=g) * 0.5 + 0.08
        t[1] = t[0]                       # two labels in the same cell with the same anchors: duplicate indices
        t[1, 4:6] *= 1.1
        if nt > 4:
            t[4, 0], t[4, 2:4] = t[3, 0], t[3, 2:4] + 0.001
        if seed == 5:                     # edge geometry
            t[2, 2:4] = torch.tensor([0.5, 0.25])          # exactly on cell boundaries of every grid
            t[3, 2:4] = torch.tensor([0.0, 0.0])           # image corner (cell 0, 0)
            t[5, 2:4] = torch.tensor([0.999, 0.999])       # last cell
            t[6, 4:6] = torch.tensor([0.9, 0.95])          # matches only the largest anchors
            t[7, 4:6] = torch.tensor([0.002, 0.002])       # matches no anchor at all
    return p, anchor_vecs, t


def main():
    _, ref_utils = import_reference()
    out = {}
    for name, v4, ciou, nc, gr, obj_pw, cls_pw, nt, seed in CASES:
        p, anchor_vecs, t = case_inputs(v4, nc, nt, seed)
        hyp = {"box": 3.54, "cls": 37.4, "obj": 64.3, "cls_pw": cls_pw, "obj_pw": obj_pw, "iou_t": 0.20, "fl_gamma": 0.0}
        if ciou:
            hyp["ciou"] = 1.0
        model = types.SimpleNamespace(hyp=hyp, gr=gr, nc=nc, cfg="config/kaist_dyolov4_x.cfg" if v4 else "config/kaist_yolov3.cfg",
                                      yolo_layers=[0, 1, 2],
                                      module_list=[types.SimpleNamespace(anchor_vec=a) for a in anchor_vecs])
        pr = [x.clone().requires_grad_(True) for x in p]
        loss = ref_utils.compute_loss(pr, t, model)
        total = loss["box_loss"] + loss["obj_loss"] + loss["class_loss"]
        total.backward()
        tcls, tbox, indices, anch = ref_utils.build_targets(p, t, model)
        out[f"{name}/meta"] = np.array([v4, ciou, nc, gr, obj_pw, cls_pw, nt], dtype=np.float64)
        out[f"{name}/targets"] = t.numpy()
        for i in range(3):
            out[f"{name}/p{i}"] = p[i].numpy()
            out[f"{name}/anchor{i}"] = anchor_vecs[i].numpy()
            out[f"{name}/grad{i}"] = (pr[i].grad if pr[i].grad is not None else torch.zeros_like(p[i])).numpy()
            b, a, gj, gi = indices[i]
            a = a if torch.is_tensor(a) else torch.zeros(0, dtype=torch.long)
            out[f"{name}/idx{i}"] = torch.stack([b, a, gj, gi], 1).numpy().astype(np.int64) if len(b) else np.zeros((0, 4), np.int64)
            out[f"{name}/tbox{i}"] = tbox[i].numpy()
            out[f"{name}/tcls{i}"] = tcls[i].numpy().astype(np.int64)
        out[f"{name}/losses"] = np.array([float(loss["box_loss"]), float(loss["obj_loss"]), float(loss["class_loss"])])
        print(name, out[f"{name}/losses"], [len(indices[i][0]) for i in range(3)])
    np.savez_compressed(HERE / "loss_cases.npz", **out)


if __name__ == "__main__":
    main()
