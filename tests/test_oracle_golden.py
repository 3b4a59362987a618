"""Pins the oracle (oracle/) to the golden vectors produced by the real reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import darknet_ref as dr
from oracle import nms_ref
from oracle import weights as ow
from dyk import cfg_zoo

MODEL_CFGS = sorted(cfg_zoo.ZOO)


def _frames(dual, B, H, W, seed=7):
    g = torch.Generator().manual_seed(seed)
    v = torch.rand((B, 3, H, W), generator=g)
    l = torch.rand((B, 3, H, W), generator=g) if dual else None
    return v, l


@pytest.mark.parametrize("name", MODEL_CFGS)
def test_oracle_model_matches_reference_outputs(name, golden_dir):
    gold = np.load(golden_dir / (name[:-4] + ".npz"))
    ref = dr.DarknetRef(cfg_zoo.materialize(name))
    st = ow.make_calibrated_state(ref, seed=0)
    # calibrated BN statistics come from the oracle's own train-mode forward
    rm = np.array([float(st[k].double().sum()) for k in sorted(st) if k.endswith("running_mean")])
    rv = np.array([float(st[k].double().sum()) for k in sorted(st) if k.endswith("running_var")])
    np.testing.assert_allclose(rm, gold["bn_mean_sums"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(rv, gold["bn_var_sums"], rtol=1e-4, atol=1e-4)
    dual = "second_index" in ref.net
    v, l = _frames(dual, int(gold["B"]), int(gold["H"]), int(gold["W"]))
    with torch.no_grad():
        (io, p), every = ref.forward(st, v, l, keep_layers=True)
        ptrain = ref.forward(st, v, l, training=True)
    # fp32 tolerance: oneDNN may pick a different summation order for another thread count
    np.testing.assert_allclose(io.numpy(), gold["io"], rtol=1e-4, atol=1e-4)
    for i, t in enumerate(p):
        np.testing.assert_allclose(t.numpy(), gold[f"p{i}"], rtol=1e-4, atol=1e-4)
    for i, t in enumerate(ptrain):
        np.testing.assert_allclose(t.numpy(), gold[f"ptrain{i}"], rtol=1e-3, atol=1e-3)
    # the input actually reaches the output (guards against the default-init collapse)
    v2, l2 = _frames(dual, int(gold["B"]), int(gold["H"]), int(gold["W"]), seed=8)
    with torch.no_grad():
        io2, _ = ref.forward(st, v2, l2)
    assert float((io2 - io).abs().max()) > 1.0
    absmean = np.array([float(e.abs().mean()) for e in every])
    yolo_rows = [i for i, d in enumerate(ref.defs) if d["type"] == "yolo"]
    keep = [i for i in range(len(absmean)) if i not in yolo_rows]
    np.testing.assert_allclose(absmean[keep], gold["layer_absmean"][keep], rtol=1e-3, atol=1e-4)


def _nms_case_names(z):
    return sorted({k.split("/")[0] for k in z.files})


def test_oracle_nms_bit_exact_against_reference(golden_dir):
    z = np.load(golden_dir / "nms_cases.npz")
    names = _nms_case_names(z)
    assert len(names) >= 10
    for c in names:
        kw = {k.split("/")[2]: z[k].tolist() for k in z.files if k.startswith(c + "/kw/")}
        res = nms_ref.non_max_suppression(z[c + "/pred"], **kw)
        counts = [-1 if r is None else r.shape[0] for r in res]
        assert counts == z[c + "/count"].tolist(), c
        for i, r in enumerate(res):
            if r is not None:
                assert np.array_equal(r, z[f"{c}/out{i}"]), (c, i)


def test_oracle_greedy_nms_equals_torchvision():
    torchvision = pytest.importorskip("torchvision")
    g = np.random.default_rng(5)
    for n in (1, 2, 50, 700):
        xy = g.uniform(0, 200, size=(n, 2)).astype(np.float32)
        wh = g.uniform(5, 80, size=(n, 2)).astype(np.float32)
        boxes = np.concatenate([xy, xy + wh], 1)
        scores = np.round(g.uniform(0, 1, size=n), 2).astype(np.float32)  # many ties
        for thr in (0.3, 0.6):
            want = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
            got = nms_ref.greedy_nms(boxes, scores, thr)
            assert np.array_equal(got, want)
