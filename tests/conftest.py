import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# The longest-running modules of the -m gpu suite (reference fixtures, compiled host, full-size properties) run last, so that under
# `pytest -x` a failure there cannot hide the quick parity tests.
_RUN_LAST = ["test_reference_deposition", "test_reference_shapefunction", "test_zz_host_cpp", "test_zz_gpu_properties",
             "test_reference_push", "test_reference_tracking"]


def pytest_collection_modifyitems(config, items):
    def rank(it):
        name = os.path.splitext(os.path.basename(str(getattr(it, "path", None) or it.fspath)))[0]
        return _RUN_LAST.index(name) + 1 if name in _RUN_LAST else 0
    items.sort(key=rank)                 # stable: the order inside a module and among the other modules is kept
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
