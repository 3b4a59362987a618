"""Native compute_loss / build_targets (csrc/yolo_loss.cu) on a B200 against
  (1) the committed outputs of the REAL reference (tests/golden/loss_cases.npz): matched targets bit-exact, losses to 1e-5
      relative, gradients with respect to every head tensor to 2e-4 of their largest magnitude (fp32 arithmetic in a
      different association order than autograd's);
  (2) the oracle (oracle/loss_ref.py, itself bit-exact against the reference on the golden cases) at BASELINE size:
      bs16, 512x640 grids, 60 labels."""
import types
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = Path(__file__).resolve().parent / "golden" / "loss_cases.npz"
CASES = ["v3_giou", "v4_ciou", "v3_ciou_nc3", "v4_giou_empty", "v3_ciou_edges", "v4_giou_nc2_pw"]


def _model(c):
    return types.SimpleNamespace(hyp=c["hyp"], gr=c["gr"], nc=c["nc"], cfg="cfg/kaist_dyolov4_x.cfg" if c["v4"] else "cfg/kaist_yolov3.cfg",
                                 yolo_layers=[0, 1, 2],
                                 module_list=[types.SimpleNamespace(anchor_vec=a) for a in c["anchors"]])


def _check(got, want, rel, what):
    scale = max(float(np.abs(want).max()), 1e-12)
    err = float(np.abs(got - want).max())
    assert err <= rel * scale, (what, err, scale)


@pytest.mark.parametrize("name", CASES)
def test_loss_matches_reference_golden(native_lib, name):
    from build_utils.utils import build_targets, compute_loss
    from test_oracle_loss import load_case
    z = np.load(GOLD)
    c = load_case(z, name)
    model = _model(c)
    p = [t.to(DEV).requires_grad_(True) for t in c["p"]]
    targets = c["targets"].to(DEV)
    tcls, tbox, indices, anch = build_targets(p, targets, model)
    for i in range(3):
        b, a, gj, gi = indices[i]
        idx = torch.stack([b, a, gj, gi], 1).cpu().numpy()
        assert np.array_equal(idx, z[f"{name}/idx{i}"]), "matched targets must be identical and in the reference's order"
        assert np.array_equal(tcls[i].cpu().numpy(), z[f"{name}/tcls{i}"])
        assert np.array_equal(tbox[i].cpu().numpy(), z[f"{name}/tbox{i}"])
        assert torch.equal(anch[i].cpu(), c["anchors"][i][a.cpu()])
    loss = compute_loss(p, targets, model)
    got = np.array([float(loss["box_loss"]), float(loss["obj_loss"]), float(loss["class_loss"])])
    want = z[f"{name}/losses"]
    assert np.allclose(got, want, rtol=1e-5, atol=1e-7), (got, want)
    (loss["box_loss"] + loss["obj_loss"] + loss["class_loss"]).backward()
    for i in range(3):
        _check(p[i].grad.cpu().numpy(), z[f"{name}/grad{i}"], 2e-4, f"{name}: grad of head {i}")


def test_loss_upstream_scaling_and_determinism(native_lib):
    """GradScaler-style scaling of the loss reaches the gradients; separate weights per loss output are honoured; two runs
    are bit-identical."""
    from build_utils.utils import compute_loss
    from test_oracle_loss import load_case
    c = load_case(np.load(GOLD), "v3_ciou_nc3")
    model = _model(c)
    targets = c["targets"].to(DEV)

    def run(wb, wo, wc):
        p = [t.to(DEV).requires_grad_(True) for t in c["p"]]
        loss = compute_loss(p, targets, model)
        (wb * loss["box_loss"] + wo * loss["obj_loss"] + wc * loss["class_loss"]).backward()
        return [t.grad.clone() for t in p]

    g1, g2 = run(1.0, 1.0, 1.0), run(1.0, 1.0, 1.0)
    assert all(torch.equal(a, b) for a, b in zip(g1, g2))
    g3 = run(1024.0, 0.0, 2.0)
    for a, b in zip(g1, g3):
        assert torch.allclose(b[..., :4], a[..., :4] * 1024.0, rtol=1e-6, atol=0)
        assert float(b[..., 4].abs().max()) == 0.0
        assert torch.allclose(b[..., 5:], a[..., 5:] * 2.0, rtol=1e-6, atol=0)


@pytest.mark.parametrize("v4,ciou", [(False, False), (True, True)])
def test_loss_baseline_size_against_oracle(native_lib, v4, ciou):
    from build_utils.utils import compute_loss
    from oracle import loss_ref
    g = torch.Generator().manual_seed(21)
    B, nt, nc = 16, 60, 1
    strides = [8, 16, 32] if v4 else [32, 16, 8]
    anchors_px = torch.tensor([[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]],
                              dtype=torch.float32)
    masks = [[0, 1, 2], [3, 4, 5], [6, 7, 8]] if v4 else [[6, 7, 8], [3, 4, 5], [0, 1, 2]]
    anchors = [anchors_px[m] / s for m, s in zip(masks, strides)]
    p = [torch.randn((B, 3, 512 // s, 640 // s, 5 + nc), generator=g) for s in strides]
    t = torch.zeros((nt, 6))
    t[:, 0] = torch.randint(0, B, (nt,), generator=g).float()
    t[:, 2:4] = torch.rand((nt, 2), generator=g) * 0.9 + 0.05
    t[:, 4] = torch.rand((nt,), generator=g) * 0.1 + 0.02
    t[:, 5] = torch.rand((nt,), generator=g) * 0.3 + 0.05
    hyp = {"box": 3.54, "cls": 37.4, "obj": 64.3, "cls_pw": 1.0, "obj_pw": 1.0, "iou_t": 0.20, "fl_gamma": 0.0}
    if ciou:
        hyp["ciou"] = 1.0
    c = dict(hyp=hyp, gr=1.0, nc=nc, v4=v4, anchors=anchors)
    po = [x.clone().requires_grad_(True) for x in p]
    lb, lo, lc = loss_ref.compute_loss(po, t, anchors, hyp, 1.0, nc, v4)
    (lb + lo + lc).backward()
    pn = [x.to(DEV).requires_grad_(True) for x in p]
    loss = compute_loss(pn, t.to(DEV), _model(c))
    (loss["box_loss"] + loss["obj_loss"] + loss["class_loss"]).backward()
    assert abs(float(loss["box_loss"]) - float(lb)) <= 1e-5 * abs(float(lb))
    assert abs(float(loss["obj_loss"]) - float(lo)) <= 1e-5 * abs(float(lo))
    for i in range(3):
        _check(pn[i].grad.cpu().numpy(), po[i].grad.numpy(), 2e-4, f"grad of head {i}")


def test_loss_bad_targets_raise_at_next_call(native_lib):
    from build_utils.utils import compute_loss
    from dyk import loss as L
    from test_oracle_loss import load_case
    c = load_case(np.load(GOLD), "v3_giou")
    model = _model(c)
    p = [t.to(DEV) for t in c["p"]]
    bad = c["targets"].clone()
    bad[0, 0] = 7          # image index outside the batch of 2: the reference raises IndexError while indexing
    compute_loss(p, bad.to(DEV), model)
    torch.cuda.synchronize()
    with pytest.raises(IndexError):
        compute_loss(p, c["targets"].to(DEV), model)
    L._check_pending(block=True)     # drain
