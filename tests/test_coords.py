"""scale_coords / clip_coords: the numpy oracle against outputs of the real reference (CPU), and the native kernel against
both on a B200 — bit-exact (fp32 subtract, IEEE divide, clamp)."""
from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).resolve().parent / "golden" / "coords_cases.npz"
CASES = ["letterbox_none", "kaist_ratio_pad", "upscaled", "odd_gain", "empty"]


def _case(z, name):
    h1, w1, h0, w0, gain, px, py = z[f"{name}/meta"].tolist()
    rp = None if gain < 0 else ((gain, gain), (px, py))
    return (int(h1), int(w1)), (int(h0), int(w0)), rp


@pytest.mark.parametrize("name", CASES)
def test_coords_oracle_matches_reference(name):
    from oracle import coords_ref
    z = np.load(GOLD)
    s1, s0, rp = _case(z, name)
    assert np.array_equal(coords_ref.scale_coords(s1, z[f"{name}/in"][:, :4], s0, rp), z[f"{name}/scaled"])
    assert np.array_equal(coords_ref.clip_coords(z[f"{name}/in"], s0), z[f"{name}/clipped"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_coords_native_matches_reference(native_lib, name):
    from build_utils.utils import clip_coords, scale_coords
    z = np.load(GOLD)
    s1, s0, rp = _case(z, name)
    pred = torch.from_numpy(z[f"{name}/in"]).cuda()
    keep = pred.clone()
    out = scale_coords(s1, pred[:, :4], s0, rp)              # a strided view of the (n, 6) NMS rows, rescaled in place
    assert np.array_equal(out.cpu().numpy(), z[f"{name}/scaled"])
    assert torch.equal(pred[:, 4:], keep[:, 4:]), "confidence / class columns must not be touched"
    boxes = torch.from_numpy(z[f"{name}/in"]).cuda()
    clip_coords(boxes, s0)
    assert np.array_equal(boxes.cpu().numpy(), z[f"{name}/clipped"])
