import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs the compiled reference under oracle/_ref")


def _have_gpu():
    if os.environ.get("B200_TESTS_ASSUME_GPU"):   # test development only: run the CLI-level "gpu" tests over the CPU mock of the ABI
        return True
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


if os.environ.get("B200_TESTS_ASSUME_GPU"):   # test development only: the Python host loads the CPU mock of the ABI as well
    try:
        from cmdstan_b200 import _build as _b
        _mock = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul", "mock", "libb200nuts.so")
        if os.path.exists(_mock):
            _b.LIB = _mock
            _b.needs_build = lambda: False
    except Exception:
        pass


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
