import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_missing():
    """Why GPU tests cannot run here (None when they must run): only "the library is there but the box has no CUDA
    device" skips them.  A missing or broken library is NOT a reason to skip: the product has no CPU fallback, so on
    a GPU box the tests then fail loudly at their first call."""
    try:
        import ctypes
        from softwarerenderer_b200 import _lib
        lib = _lib.load()
    except Exception:
        return None
    ctx = ctypes.c_void_p()
    rc = lib.swr_create(ctypes.byref(ctx), 0)
    if rc == 0:
        lib.swr_destroy(ctx)
        return None
    return lib.swr_last_error().decode() if rc == -20 else None


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU: skip the gpu-marked tests instead of failing at the first one."""
    if not any("gpu" in item.keywords for item in items):
        return
    why = _gpu_missing()
    if why is None:
        return
    skip = pytest.mark.skip(reason=f"needs a CUDA device: {why}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU checkers (test infrastructure): builds oracle/libswr_oracle.so on demand."""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
