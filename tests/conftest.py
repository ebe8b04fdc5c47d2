import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import ctypes
            cudart = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
        except OSError:
            return False
    n = ctypes.c_int(0)
    return cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Build libnbody_b200.so and the oracle if they are not there yet."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def nb(built):
    import mini_nbody_b200
    return mini_nbody_b200


@pytest.fixture(scope="session")
def orc(built):
    import oracle_lib
    oracle_lib.load("parity")
    return oracle_lib
