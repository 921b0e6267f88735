import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# no CLIP checkpoint exists offline: the towers under test are seeded random weights (an explicit opt-in, see clip_compat.load)
os.environ.setdefault("CLIPPPO_ALLOW_RANDOM_WEIGHTS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def native():
    """The C-ABI library, built if needed.  GPU tests must run on it - never on a fallback."""
    from clip_ppo_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_clipppo_build", os.path.join(ROOT, "clip-ppo_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return _native.lib()
