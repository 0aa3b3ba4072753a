import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # TSGPU_TEST_DOUBLE=1 (set only by tests/test_gpu_tests_dryrun.py, together with TSGPU_LIB_PATH pointing at the
    # oracle-backed test double of the C-ABI) lets the gpu-marked tests execute their Python side on a machine without a
    # GPU, so a typo in a test that cannot run here does not first show up on the GPU box.
    if _has_gpu() or os.environ.get("TSGPU_TEST_DOUBLE") == "1":
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
