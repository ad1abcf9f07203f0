import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) and the built libmerv_fusion.so")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without CUDA must fail loudly rather than skip silently: only auto-skip when the
    # user did not ask for the gpu marker explicitly.
    import torch

    if torch.cuda.is_available() or "gpu" in (config.getoption("-m") or ""):
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
