"""Batched NMS kernel: bit-exact against the golden vectors of the real reference and against the oracle on
larger seeded inputs (integer/index work -> exact equality)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(pred, **kw):
    from build_utils.utils import non_max_suppression
    out = non_max_suppression(torch.from_numpy(pred).to(DEV), **kw)
    return [None if o is None else o.cpu().numpy() for o in out]


def _assert_same(got, want, what):
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        if w is None:
            assert g is None, (what, i)
        else:
            assert g is not None and g.shape == w.shape, (what, i, None if g is None else g.shape, w.shape)
            assert np.array_equal(g, w), (what, i, float(np.abs(g - w).max()))


def test_nms_golden_cases_bit_exact(native_lib, golden_dir):
    z = np.load(golden_dir / "nms_cases.npz")
    for c in sorted({k.split("/")[0] for k in z.files}):
        kw = {k.split("/")[2]: z[k].tolist() for k in z.files if k.startswith(c + "/kw/")}
        counts = z[c + "/count"].tolist()
        want = [None if n < 0 else z[f"{c}/out{i}"] for i, n in enumerate(counts)]
        _assert_same(_run(z[c + "/pred"], **kw), want, c)


@pytest.mark.parametrize("B,rows,nc,regime", [(1, 20160, 1, "dense"), (4, 20160, 1, "sparse"), (16, 5040, 1, "dense"),
                                              (2, 3000, 4, "sparse"), (3, 1, 1, "dense"), (2, 4097, 2, "dense")])
def test_nms_matches_oracle(native_lib, B, rows, nc, regime):
    from oracle import nms_ref
    g = np.random.default_rng(B * 1000 + rows)
    pred = np.zeros((B, rows, 5 + nc), dtype=np.float32)
    centres = g.uniform(40, 600, size=(B, 8, 2))
    which = g.integers(0, 8, size=(B, rows))
    pred[:, :, 0:2] = np.take_along_axis(centres, which[:, :, None].repeat(2, 2), 1) + g.normal(0, 4, size=(B, rows, 2))
    pred[:, :, 2] = g.uniform(1, 60, size=(B, rows))
    pred[:, :, 3] = g.uniform(1, 120, size=(B, rows))
    mu = 3.0 if regime == "dense" else -6.0
    pred[:, :, 4] = 1 / (1 + np.exp(-g.normal(mu, 2, size=(B, rows))))
    pred[:, :, 5:] = np.round(g.uniform(0.2, 1.0, size=(B, rows, nc)), 2)
    pred = pred.astype(np.float32)
    for kw in (dict(conf_thres=0.001, iou_thres=0.6, multi_label=False), dict(conf_thres=0.1, iou_thres=0.5, multi_label=True)):
        _assert_same(_run(pred, **kw), nms_ref.non_max_suppression(pred, **kw), (B, rows, nc, regime, kw))


def test_nms_does_not_mutate_input_and_handles_empty_batch(native_lib):
    from build_utils.utils import non_max_suppression
    pred = torch.rand((2, 100, 6), device=DEV)
    pred[..., 2:4] *= 50
    keep = pred.clone()
    non_max_suppression(pred, 0.1, 0.6, multi_label=False)
    assert torch.equal(pred, keep)
    assert non_max_suppression(pred[:0]) == []
