"""Drop-in boundary checks that need no GPU: the C-ABI library loads and exports exactly what the header
declares, the Python API mirrors the reference's module tree, and the product never touches the oracle."""
import re
from pathlib import Path

import pytest
import torch

from conftest import PKG, REPO
from dyk import _native, cfg_zoo


def _header_symbols():
    text = (REPO / "include" / "dyk_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dyk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native_lib):
    syms = _header_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(native_lib, s), f"{s} declared in include/dyk_b200.h but not exported"
    assert sorted(_native.SIGNATURES) == syms, "ctypes prototypes and header went out of sync"
    assert native_lib.dyk_abi_version() == 7
    import ctypes
    from dyk import _native as nat_mod
    assert native_lib.dyk_conv_params_size() == ctypes.sizeof(nat_mod.ConvParams)      # the ctypes mirror has the C layout


def test_library_is_sm100a_tcgen05(native_lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(_native.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass, "no tcgen05.mma in the conv kernel"
    assert "UTMALDG" in sass and "UTMASTG" in sass, "no TMA loads/stores"
    assert "LDTM" in sass, "no tcgen05.ld epilogue"


def test_product_never_imports_oracle_or_reference():
    for f in PKG.rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
        assert "/root/reference" not in src, f


def test_cpu_tensors_fail_loudly():
    import models
    m = models.YOLO(cfg_zoo.materialize("kaist_yolov3.cfg"), (64, 96)).eval()
    with pytest.raises(_native.NativeError):
        m(torch.rand(1, 3, 64, 96))
    from build_utils.utils import non_max_suppression
    with pytest.raises(_native.NativeError):
        non_max_suppression(torch.rand(1, 10, 6))


@pytest.mark.parametrize("name", sorted(cfg_zoo.ZOO))
def test_state_dict_layout_matches_reference_names(name):
    """state_dict keys/shapes are those the oracle derives from the reference's create_modules."""
    import models
    from oracle import darknet_ref as dr
    path = cfg_zoo.materialize(name)
    m = models.YOLO(path, (64, 96))
    ref = dr.DarknetRef(path)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.shapes.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref.shapes[k]), k
    assert m.yolo_layers == [i for i, d in enumerate(m.module_defs) if d["type"] == "yolo"]
    for i in m.yolo_layers:
        head = m.module_list[i - 1][0]
        b = head.bias.detach().view(3, -1)
        assert torch.allclose(b[:, 4].mean(), torch.tensor(-4.5), atol=0.2)  # smart bias prior (models.py:139-142)
    names = {type(x).__name__ for x in m.module_list}
    assert {"FeatureConcat", "WeightedFeatureFusion", "YOLOLayer", "Sequential"} <= names


def test_plan_fuses_and_places(native_lib):
    import models
    from dyk import plan
    m = models.YOLO(cfg_zoo.materialize("kaist_dyolov4_fshare_global_concat_se3.cfg"), (512, 640)).eval()
    raw, vals, _, _ = plan.build_ops(m, 512, 640, True)
    ops_ = plan.fuse(raw)
    plan.mark_heads(ops_)
    plan.place_concats(ops_)
    plan.liveness(ops_)
    kinds = [o.kind for o in ops_]
    assert kinds.count("upsample") == 0 and kinds.count("yolo") == 3
    assert sum(1 for o in ops_ if o.kind == "conv" and o.res is not None) == 46
    assert sum(len(o.copies) for o in ops_ if o.kind == "concat") == 0      # every concat is zero-copy
    assert sum(1 for o in ops_ if o.kind == "conv" and o.out_f32) == 3
    assert len(vals) == len(m.module_defs)
    with pytest.raises(ValueError):
        plan.build_ops(m, 512, 640, False)   # dual cfg called with one modality


def test_ctypes_prototypes_have_the_header_arity():
    text = (REPO / "include" / "dyk_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"\b(dyk_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S)
    assert len(protos) >= 19
    for name, args in protos:
        args = args.strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        assert len(_native.SIGNATURES[name][1]) == n, f"{name}: header has {n} parameters"
    # struct dyk_conv_params field count
    body = re.search(r"typedef struct dyk_conv_params \{(.*?)\} dyk_conv_params;", text, flags=re.S).group(1)
    fields = [f for stmt in body.split(";") for f in stmt.split(",") if f.strip()]
    assert len(fields) == len(_native.ConvParams._fields_)
