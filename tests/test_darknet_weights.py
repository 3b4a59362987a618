"""load_darknet_weights (reference models.py:318-364): a Darknet .weights blob written in the documented order
(header int32 x3 + int64 `seen`; per [convolutional]: BN bias, weight, running_mean, running_var | conv bias; conv weight)
must land in the right tensors, bump their version counters (so cached native plans re-pack) and honour `cutoff`."""
import numpy as np
import pytest
import torch

import models
from dyk import cfg_zoo
from oracle import darknet_ref as dr
from oracle import weights as ow


def _write_weights(path, defs, state, cutoff=None):
    with open(path, "wb") as f:
        np.array([0, 2, 5], dtype=np.int32).tofile(f)
        np.array([12345], dtype=np.int64).tofile(f)
        for i, d in enumerate(defs[:cutoff]):
            if d["type"] != "convolutional":
                continue
            pre = f"module_list.{i}."
            if d["batch_normalize"]:
                for k in ("BatchNorm2d.bias", "BatchNorm2d.weight", "BatchNorm2d.running_mean", "BatchNorm2d.running_var"):
                    state[pre + k].numpy().astype(np.float32).tofile(f)
            else:
                state[pre + "Conv2d.bias"].numpy().astype(np.float32).tofile(f)
            state[pre + "Conv2d.weight"].numpy().astype(np.float32).tofile(f)


@pytest.mark.parametrize("cfg", ["kaist_yolov3.cfg", "kaist_dyolov3_add_sl.cfg"])
def test_round_trip(tmp_path, cfg):
    path = cfg_zoo.materialize(cfg)
    ref = dr.DarknetRef(path)
    want = ow.make_calibrated_state(ref, seed=3)
    blob = str(tmp_path / "m.weights")
    model = models.YOLO(path, (128, 160))
    _write_weights(blob, model.module_defs, want)
    versions = {k: v._version for k, v in model.state_dict(keep_vars=True).items()}
    models.load_darknet_weights(model, blob)
    got = model.state_dict()
    assert list(model.version) == [0, 2, 5] and int(model.seen[0]) == 12345
    touched = 0
    for k, v in want.items():
        if k.endswith("num_batches_tracked") or ".w" == k[-2:]:
            continue
        if "Conv2d" in k or "BatchNorm2d" in k:
            assert torch.equal(got[k], v), k
            touched += 1
    assert touched > 100
    bumped = [k for k, v in model.state_dict(keep_vars=True).items() if v._version != versions[k]]
    assert len(bumped) >= touched      # every loaded tensor's version counter moved: stale native plans are detected


def test_cutoff_and_short_file(tmp_path):
    path = cfg_zoo.materialize("kaist_yolov3.cfg")
    ref = dr.DarknetRef(path)
    want = ow.make_calibrated_state(ref, seed=4)
    model = models.YOLO(path, (128, 160))
    before = {k: v.clone() for k, v in model.state_dict().items()}
    blob = str(tmp_path / "backbone.weights")
    _write_weights(blob, model.module_defs, want, cutoff=75)           # darknet53.conv.74-style backbone file
    models.load_darknet_weights(model, blob, cutoff=75)
    got = model.state_dict()
    assert torch.equal(got["module_list.73.Conv2d.weight"], want["module_list.73.Conv2d.weight"])
    assert torch.equal(got["module_list.75.Conv2d.weight"], before["module_list.75.Conv2d.weight"])   # untouched beyond cutoff
    with pytest.raises(ValueError):
        models.load_darknet_weights(model, blob)                      # whole model from a backbone-only file
