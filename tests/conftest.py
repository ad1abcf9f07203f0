import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def _ensure_library_built():
    """The CUDA library is a build artefact (git-ignored): (re)build it with nvcc when this checkout has none or its sources
    changed (digest check, a no-op otherwise).  Same step as __graft_entry__.build(); still no fallback if nvcc is missing."""
    import importlib.util
    import shutil

    lib = os.path.join(REPO, "merv_b200", "libmerv_fusion.so")
    if os.path.isfile(lib) and shutil.which("nvcc") is None and not os.path.isfile("/usr/local/cuda/bin/nvcc"):
        return  # a box without the toolchain uses the library that travelled with the snapshot

    spec = importlib.util.spec_from_file_location("merv_b200_build", os.path.join(REPO, "merv_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) and the built libmerv_fusion.so")
    _ensure_library_built()


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
