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


# GPU tests written after the last hardware session of the round (profiles/README.md, "validated on hardware"): they run
# last, so that a `-x` run reports the hardware-validated suite before it reaches them.  Empty this list once they have run.
NOT_YET_RUN_ON_HARDWARE = (
    "test_static_hmc_matched_stream_vs_oracle",
    "test_fp64_fma_peak_is_measurable",
    "test_cli_log_prob_method_vs_reference",
    "test_cli_diagnostic_file_vs_reference",
    "test_diagnostic_rows_vs_oracle",   # its content ran on hardware as tools/gpu_quickcheck.py (profiles/r01l_quickcheck.txt), the test itself did not
)


def pytest_collection_modifyitems(config, items):
    late = [it for it in items if it.name.split("[")[0] in NOT_YET_RUN_ON_HARDWARE]
    if late:
        items[:] = [it for it in items if it not in late] + late
    if _have_gpu():
        # a kernel that never returns must not hold the box until the caller's limit: every GPU test gets a wall-clock limit
        # (pytest-timeout, thread method: the process is ended, which also releases the CUDA context)
        if config.pluginmanager.hasplugin("timeout"):
            for it in items:
                if "gpu" in it.keywords and not it.get_closest_marker("timeout"):
                    it.add_marker(pytest.mark.timeout(900, method="thread"))
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
