import os
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
PKG = REPO / "double-yolo-kaist_b200"
for p in (str(REPO), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE = Path(os.environ.get("DYK_REFERENCE", "/root/reference"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _b200_available() -> bool:
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:  # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a machine without an sm_100 device."""
    if _b200_available():
        return
    skip = pytest.mark.skip(reason="needs a B200 (sm_100) GPU")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return REPO / "tests" / "golden"


@pytest.fixture(scope="session")
def native_lib():
    """Builds (if stale) and loads libdyk_b200.so; no GPU needed for loading."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("dyk_build", PKG / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from dyk import _native
    return _native.load()
