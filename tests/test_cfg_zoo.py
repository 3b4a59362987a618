"""The generated cfgs must describe exactly the reference's models (parse result equality)."""
import importlib.util
import json

import numpy as np
import pytest

from conftest import REFERENCE, REPO
from build_utils.parse_config import parse_model_cfg, parse_model_cfg_text
from dyk import cfg_zoo

DIGESTS = json.loads((REPO / "tests" / "golden" / "cfg_digests.json").read_text())


def _same(a, b):
    if set(a) != set(b):
        return False
    for k in a:
        if isinstance(a[k], np.ndarray):
            if not np.array_equal(a[k], b[k]):
                return False
        elif a[k] != b[k] or type(a[k]) is not type(b[k]):
            return False
    return True


@pytest.mark.parametrize("name", sorted(cfg_zoo.ZOO))
def test_generated_cfg_matches_committed_digest(name):
    blocks = parse_model_cfg(cfg_zoo.materialize(name))
    assert cfg_zoo.structural_digest(blocks) == DIGESTS[name]


@pytest.mark.skipif(not (REFERENCE / "config").exists(), reason="reference tree not present")
@pytest.mark.parametrize("name", sorted(cfg_zoo.ZOO))
def test_generated_cfg_equals_reference_file(name):
    spec = importlib.util.spec_from_file_location("ref_parse_config", REFERENCE / "build_utils" / "parse_config.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    theirs = ref.parse_model_cfg(str(REFERENCE / "config" / name))
    mine = parse_model_cfg(cfg_zoo.materialize(name))
    assert len(mine) == len(theirs)
    for i, (a, b) in enumerate(zip(mine, theirs)):
        assert _same(a, b), f"block {i - 1}: {a} != {b}"


@pytest.mark.skipif(not (REFERENCE / "config").exists(), reason="reference tree not present")
def test_parser_equals_reference_parser_on_every_shipped_cfg():
    spec = importlib.util.spec_from_file_location("ref_parse_config", REFERENCE / "build_utils" / "parse_config.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    files = sorted((REFERENCE / "config").glob("*.cfg"))
    assert len(files) == 28
    for f in files:
        a, b = parse_model_cfg(str(f)), ref.parse_model_cfg(str(f))
        assert len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b)), f.name


def test_parser_dialect():
    blocks = parse_model_cfg_text("[net]\n;width=608\nhue = .1\nbatch=64\n# c\n\n[convolutional]\nfilters = 18\nsize=1\n"
                                  "stride=1\npad=1\nactivation=linear\n[yolo]\nmask = 0,1\nanchors = 1,2, 3,4\n"
                                  "ignore_thresh = .7\nclasses=1\n")
    assert blocks[0] == {"type": "net", ";width": 608, "hue": ".1", "batch": 64}
    assert blocks[1]["batch_normalize"] == 0 and blocks[1]["filters"] == 18
    assert blocks[2]["mask"] == [0, 1] and blocks[2]["ignore_thresh"] == ".7"
    assert blocks[2]["anchors"].shape == (2, 2) and blocks[2]["anchors"].dtype == np.float64
    with pytest.raises(ValueError):
        parse_model_cfg_text("[net]\n[convolutional]\nbogus_key=1\n")
    with pytest.raises(FileNotFoundError):
        parse_model_cfg("/nonexistent/x.cfg")
    with pytest.raises(FileNotFoundError):
        parse_model_cfg(__file__)  # exists but is not a .cfg
