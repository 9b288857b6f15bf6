import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        from lib import _native
        return _native.lib().bf_device_count() > 0
    except Exception:  # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU should skip rather than fail; `-m "not gpu"` never loads CUDA.
    if config.getoption("-m") and "not gpu" in config.getoption("-m"):
        return
    gpu_items = [i for i in items if "gpu" in i.keywords]
    if gpu_items and not _has_gpu():
        skip = pytest.mark.skip(reason="no CUDA device visible")
        for i in gpu_items:
            i.add_marker(skip)
