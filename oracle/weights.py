"""ORACLE — TEST INFRASTRUCTURE ONLY.  Seeded, *calibrated* synthetic weights.

No trained checkpoint ships with the reference (README.md:153-164 are external links) and PyTorch's default
initialisation makes eval-mode outputs input-independent (activations decay to exactly 0 by layer ~144,
SURVEY.md Appendix E), so a parity test on default weights passes vacuously.  The recipe used by every
test, by smoke() and by bench.py:

  1. fill every tensor of the state dict from one seeded CPU generator, by (sorted) state_dict name;
  2. calibrate BatchNorm running statistics as a cumulative average over two train-mode forwards on seeded
     frames (nn.BatchNorm2d(momentum=None) semantics), so activations keep O(1) scale at every depth.

make_golden.py applies the same recipe to the real reference model; `forward_train` below is the oracle's
own train-mode forward, so agreement of the calibrated statistics is itself a parity check of train-mode BN.
"""
from __future__ import annotations

import math

import torch


def fill_state(shapes: dict, seed: int = 0, nc: int = 1) -> dict:
    """name -> fp32 CPU tensor for every entry of `shapes` (an OrderedDict name -> shape)."""
    g = torch.Generator().manual_seed(seed)
    st = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        leaf = name.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.long)
        elif leaf == "running_mean":
            t = torch.zeros(shape)
        elif leaf == "running_var":
            t = torch.ones(shape)
        elif leaf == "w":                                   # WeightedFeatureFusion.w
            t = torch.randn(shape, generator=g) * 0.5
        elif len(shape) == 4:                               # conv weight
            fan_in = shape[1] * shape[2] * shape[3]
            t = (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(3.0 / fan_in)
        elif ".BatchNorm2d." in name or ".conv.1." in name or ".conv.4." in name:
            t = torch.rand(shape, generator=g) + 0.5 if leaf == "weight" else torch.randn(shape, generator=g) * 0.1
        else:                                               # conv / SE biases
            t = torch.randn(shape, generator=g) * 0.1
            if ".Conv2d.bias" in name:                      # detection head: models.py:139-142 prior
                b = t.view(-1, nc + 5)
                b[:, 4] += -4.5
                b[:, 5:] += math.log(0.6 / (nc - 0.99))
        st[name] = t
    return st


def calibration_frames(dual: bool, h: int = 128, w: int = 160, n: int = 2, batches: int = 2, seed: int = 1234):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(batches):
        v = torch.rand((n, 3, h, w), generator=g)
        l = torch.rand((n, 3, h, w), generator=g) if dual else None
        out.append((v, l))
    return out


def calibrate(ref, st: dict, frames=None) -> dict:
    """Runs the oracle's train-mode forward with cumulative-average BN statistics over `frames`."""
    dual = "second_index" in ref.net
    frames = frames or calibration_frames(dual)
    for k in st:
        if k.endswith("running_mean"):
            st[k].zero_()
        elif k.endswith("running_var"):
            st[k].fill_(1.0)
        elif k.endswith("num_batches_tracked"):
            st[k].zero_()
    with torch.no_grad():
        for v, l in frames:
            ref.forward(st, v, l, training=True, bn_momentum=None)
    return st


def make_calibrated_state(ref, seed: int = 0) -> dict:
    st = fill_state(ref.shapes, seed)
    return calibrate(ref, st)
