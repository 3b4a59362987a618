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
