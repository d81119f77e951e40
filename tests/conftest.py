import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The C-ABI libraries are built in-tree; build them if a fresh checkout has none."""
    from pampa_b200 import build
    build.build_all(force=False)


def _cuda_devices():
    """Devices the CUDA runtime sees, asked of the driver library directly (no torch import at collection)."""
    import ctypes
    try:
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cuda.cuInit(0) != 0 or cuda.cuDeviceGetCount(ctypes.byref(n)) != 0:
            return 0
        return n.value
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a device: without one they are skipped, not failed, so that a plain `pytest tests`
    on a CPU box is green (the product path itself still fails loudly without a GPU: test_no_cpu_fallback)."""
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
