"""Generates the golden fixtures in this directory from the REAL reference (run in the build container).

    python tests/golden/make_golden.py            # needs /root/reference; writes tests/golden/*.npz

The reference (Ye-zixiao/Double-YOLO-Kaist) is imported unmodified through two import shims it needs on
a current stack (`cv2.cv2`, a stub `matplotlib`).  For each model cfg the seeded
calibrated-weights recipe of oracle/weights.py is applied to the reference's own `models.YOLO`, the BN
statistics are calibrated with the reference's own train-mode forward, and its eval-mode outputs on
seeded frames are stored.  For NMS the reference's `non_max_suppression` (which calls
torchvision.ops.nms) is run on seeded predictions.  /root/reference does not exist on the GPU box, hence
the committed vectors.
"""
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path(os.environ.get("DYK_REFERENCE", "/root/reference"))

MODEL_CASES = [  # (cfg name, H, W, batch)
    ("kaist_yolov3.cfg", 64, 96, 1),
    ("kaist_dyolov3_add_sl.cfg", 64, 96, 2),
    ("kaist_dyolov4_fshare_global_concat_se3.cfg", 64, 96, 2),
    ("kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg", 64, 96, 2),
]


def import_reference():
    import cv2
    cv2.cv2 = cv2
    sys.modules["cv2.cv2"] = cv2
    mpl = types.ModuleType("matplotlib")
    mpl.rc = lambda *a, **k: None
    sys.modules.setdefault("matplotlib", mpl)
    sys.path.insert(0, str(REF))
    import models  # noqa: E402  (the reference's)
    from build_utils import utils as ref_utils  # noqa: E402
    assert str(REF) in models.__file__
    return models, ref_utils


def eval_frames(dual, B, H, W, seed=7):
    g = torch.Generator().manual_seed(seed)
    v = torch.rand((B, 3, H, W), generator=g)
    l = torch.rand((B, 3, H, W), generator=g) if dual else None
    return v, l


def golden_model(models, name, H, W, B):
    sys.path.insert(0, str(REPO))
    from oracle import weights as ow
    torch.manual_seed(0)
    model = models.YOLO(str(REF / "config" / name), img_size=(H, W))
    sd = model.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    filled = ow.fill_state(shapes, seed=0)
    model.load_state_dict(filled, strict=True)
    dual = "second_index" in model.net_info
    # calibrate BN running statistics with the reference's own train-mode forward (cumulative average)
    bns = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for bn in bns:
        bn.momentum = None
        bn.reset_running_stats()
    model.train()
    with torch.no_grad():
        for v, l in ow.calibration_frames(dual):
            model(v, l) if dual else model(v)
    for bn in bns:
        bn.momentum = 0.1
    model.eval()
    v, l = eval_frames(dual, B, H, W)
    feats = []
    hooks = [m.register_forward_hook(lambda mod, inp, out: feats.append(out)) for m in model.module_list]
    with torch.no_grad():
        io, p = model(v, l) if dual else model(v)
    for h in hooks:
        h.remove()
    layer_absmean = np.array([float(f.abs().mean()) if isinstance(f, torch.Tensor) else float(f[0].abs().mean())
                              for f in feats], dtype=np.float64)
    sd = model.state_dict()
    rm = np.array([float(sd[k].double().sum()) for k in sorted(sd) if k.endswith("running_mean")])
    rv = np.array([float(sd[k].double().sum()) for k in sorted(sd) if k.endswith("running_var")])
    # training-mode heads on the same frames (batch statistics, no running-stat side effects kept)
    model.train()
    with torch.no_grad():
        ptrain = model(v, l) if dual else model(v)
    out = {"io": io.numpy(), "layer_absmean": layer_absmean, "bn_mean_sums": rm, "bn_var_sums": rv,
           "H": H, "W": W, "B": B}
    for i, t in enumerate(p):
        out[f"p{i}"] = t.numpy()
    for i, t in enumerate(ptrain):
        out[f"ptrain{i}"] = t.numpy()
    return out


def nms_cases():
    """(name, prediction array, kwargs)"""
    cases = []
    g = np.random.default_rng(2)

    def clustered(B, rows, nc, obj_mu):
        pred = np.zeros((B, rows, 5 + nc), dtype=np.float32)
        for b in range(B):
            centres = g.uniform(40, 600, size=(8, 2))
            which = g.integers(0, 8, size=rows)
            pred[b, :, 0:2] = centres[which] + g.normal(0, 4, size=(rows, 2))
            pred[b, :, 2] = g.uniform(10, 60, size=rows)
            pred[b, :, 3] = g.uniform(20, 120, size=rows)
            pred[b, :, 4] = 1 / (1 + np.exp(-g.normal(obj_mu, 2, size=rows)))
            pred[b, :, 5:] = g.uniform(0.2, 1.0, size=(rows, nc))
        return pred.astype(np.float32)

    cases.append(("sparse_nc1", clustered(3, 600, 1, -3.0), dict(conf_thres=0.01, iou_thres=0.6, multi_label=False)))
    cases.append(("dense_nc1", clustered(2, 900, 1, 2.0), dict(conf_thres=0.001, iou_thres=0.6, multi_label=False)))
    cases.append(("multilabel_nc3", clustered(2, 400, 3, 0.0), dict(conf_thres=0.1, iou_thres=0.5, multi_label=True)))
    cases.append(("bestclass_nc3", clustered(2, 400, 3, 0.0), dict(conf_thres=0.1, iou_thres=0.5, multi_label=False)))
    cases.append(("agnostic_nc3", clustered(2, 400, 3, 0.0),
                  dict(conf_thres=0.1, iou_thres=0.5, multi_label=True, agnostic=True)))
    cases.append(("classes_nc3", clustered(2, 400, 3, 0.0),
                  dict(conf_thres=0.1, iou_thres=0.5, multi_label=True, classes=[0, 2])))
    cases.append(("maxnum5", clustered(2, 500, 1, 1.0),
                  dict(conf_thres=0.01, iou_thres=0.6, multi_label=False, max_num=5)))
    # ties: many identical scores and duplicated boxes -> exercises the stable sort tie-break
    t = clustered(2, 300, 1, 1.0)
    t[:, :, 4] = np.round(t[:, :, 4] * 8) / 8
    t[:, :, 5] = 1.0
    t[:, 100:200, :4] = t[:, 0:100, :4]
    cases.append(("ties_nc1", t, dict(conf_thres=0.01, iou_thres=0.6, multi_label=False)))
    # IoU exactly at the threshold (must NOT be suppressed: strict '>') and just above it
    e = np.zeros((1, 4, 6), dtype=np.float32)
    e[0, 0] = [100, 100, 40, 40, 0.9, 1.0]
    e[0, 1] = [120, 100, 40, 40, 0.8, 1.0]     # IoU 1/3 with box 0
    e[0, 2] = [300, 300, 40, 40, 0.7, 1.0]
    e[0, 3] = [300, 310, 40, 40, 0.6, 1.0]     # IoU 0.6 with box 2 -> 30*40/(2*1600-1200)
    cases.append(("iou_edge", e, dict(conf_thres=0.1, iou_thres=1.0 / 3.0, multi_label=False)))
    cases.append(("iou_edge06", e, dict(conf_thres=0.1, iou_thres=0.6, multi_label=False)))
    # one image with nothing above threshold -> None, boxes outside the (2, 4096) size window dropped
    z = clustered(3, 200, 1, 0.0)
    z[1, :, 4] = 0.0001
    z[2, :100, 2] = 1.5
    z[2, 100:, 3] = 5000
    cases.append(("empty_images", z, dict(conf_thres=0.01, iou_thres=0.6, multi_label=False)))
    return cases


def main():
    models, ref_utils = import_reference()
    only = sys.argv[1:]
    for name, H, W, B in MODEL_CASES:
        if only and name not in only and "models" not in only:
            continue
        if not (REF / "config" / name).exists():
            continue
        out = golden_model(models, name, H, W, B)
        np.savez_compressed(HERE / (name.replace(".cfg", "") + ".npz"), **out)
        print(name, "io", out["io"].shape, "sum", float(out["io"].astype(np.float64).sum()),
              "mean|act| min/max", out["layer_absmean"].min(), out["layer_absmean"].max())
    if not only or "nms" in only:
        pack = {}
        for cname, pred, kw in nms_cases():
            res = ref_utils.non_max_suppression(torch.from_numpy(pred.copy()), **kw)
            pack[cname + "/pred"] = pred
            for k, v in kw.items():
                pack[cname + "/kw/" + k] = np.asarray(v)
            pack[cname + "/count"] = np.array([-1 if r is None else r.shape[0] for r in res])
            for i, r in enumerate(res):
                if r is not None:
                    pack[f"{cname}/out{i}"] = r.numpy()
            print("nms", cname, pack[cname + "/count"])
        np.savez_compressed(HERE / "nms_cases.npz", **pack)


if __name__ == "__main__":
    main()
