import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): compile the CUDA extension once if the library
    is missing (nvcc cross-compiles sm_100a without a GPU).  The oracle builds itself on first use."""
    from qumcmc_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build(verbose=False)


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
