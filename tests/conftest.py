import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """Builds libmmad_b200.so (nvcc cross-compiles without a GPU) and returns its path."""
    from multimodal_ad_b200 import build

    return build.build()


@pytest.fixture(scope="session")
def c_oracle():
    """ctypes handle of the plain-C oracle (oracle/roi_oracle.c), built on demand."""
    import ctypes
    import subprocess

    so = os.path.join(ROOT, "oracle", "_build", "libroi_oracle.so")
    src = os.path.join(ROOT, "oracle", "roi_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    lib.roi_pool_oracle_c.restype = ctypes.c_int
    lib.roi_pool_oracle_c.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                      ctypes.c_int32] + [ctypes.c_void_p] * 4
    return lib
