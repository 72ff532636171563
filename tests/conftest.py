import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests skip (instead of failing on 'no NVIDIA driver') on a box without a CUDA device or library."""
    import torch
    reason = None
    if not torch.cuda.is_available():
        reason = "needs a CUDA device"
    else:
        from moco_flow_b200 import _lib
        if not os.path.exists(_lib.LIB_PATH):
            reason = "libmoco_flow_b200.so not built"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
