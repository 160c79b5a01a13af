import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        from isocon_b200 import _binding
        return _binding.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # A GPU test on a box without a GPU is an error of the run, not a silent skip -- unless the
    # caller did not ask for GPU tests explicitly (plain `pytest tests/` on the CPU container).
    if _has_gpu():
        return
    markexpr = config.getoption("-m") or ""
    if "gpu" in markexpr and "not gpu" not in markexpr:
        return  # let them fail loudly
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
