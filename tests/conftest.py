import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import ciao_pkg  # noqa: E402

ciao_pkg.load()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count():
    try:
        import ctypes
        from ciaoalgorithms_jl_b200 import _lib
        n = ctypes.c_int(0)
        return n.value if _lib.load().ciao_device_count(ctypes.byref(n)) == 0 else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a host without a GPU: the gpu-marked tests are skipped instead of failing in ciao_create
    (the product itself has no CPU fallback and says so loudly)."""
    if not any("gpu" in it.keywords for it in items):
        return
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (libciao_cuda has no CPU fallback); run with -m gpu on the B200 box")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
