"""pytest configuration: `-m gpu` tests need a B200 (they call the CUDA library through the C ABI);
everything else runs on CPU (oracle vs golden vectors, host logic, symbol export, gloo data parallel)."""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def L():
    if not os.path.exists(os.path.join(ROOT, "sg-gan-tf2_b200", "libsggan_sm100.so")):
        import __graft_entry__
        __graft_entry__.build()
    return importlib.import_module("sg-gan-tf2_b200._lib")


@pytest.fixture(scope="session")
def O():
    return importlib.import_module("sggan_oracle")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"))
