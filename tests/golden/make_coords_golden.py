"""Generates tests/golden/coords_cases.npz from the REAL reference's scale_coords / clip_coords
(build_utils/utils.py:60-92), run in the build container:  python tests/golden/make_coords_golden.py"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import import_reference  # noqa: E402

CASES = [  # name, img1 (h, w), img0 (h, w), ratio_pad or None, n, seed
    ("letterbox_none", (512, 640), (480, 640), None, 37, 1),
    ("kaist_ratio_pad", (512, 640), (512, 640), ((1.0, 1.0), (0.0, 0.0)), 100, 2),
    ("upscaled", (384, 512), (1080, 1920), ((0.26666668, 0.26666668), (0.0, 48.0)), 64, 3),
    ("odd_gain", (416, 416), (375, 500), None, 50, 4),
    ("empty", (512, 640), (480, 640), None, 0, 5),
]


def main():
    _, ref_utils = import_reference()
    out = {}
    for name, s1, s0, rp, n, seed in CASES:
        g = torch.Generator().manual_seed(seed)
        boxes = torch.rand((n, 6), generator=g) * torch.tensor([s1[1] * 1.2, s1[0] * 1.2, s1[1] * 1.2, s1[0] * 1.2, 1, 1]) - 20.0
        got = ref_utils.scale_coords(s1, boxes.clone()[:, :4], s0, rp)
        clipped = boxes.clone()
        ref_utils.clip_coords(clipped, s0)
        out[f"{name}/in"] = boxes.numpy()
        out[f"{name}/scaled"] = got.numpy()
        out[f"{name}/clipped"] = clipped.numpy()
        out[f"{name}/meta"] = np.array([s1[0], s1[1], s0[0], s0[1], -1 if rp is None else rp[0][0], 0 if rp is None else rp[1][0],
                                        0 if rp is None else rp[1][1]], dtype=np.float64)
        print(name, got.shape)
    np.savez_compressed(HERE / "coords_cases.npz", **out)


if __name__ == "__main__":
    main()
