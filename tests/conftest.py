import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_present():
    """Is there a CUDA device at all?  Asked of the driver library directly (no torch import, no dependence on our own extension:
    a missing libfinegpu.so on a GPU box must still FAIL the GPU tests, not skip them)."""
    import ctypes
    try:
        cuda = ctypes.CDLL("libcuda.so.1")
    except OSError:
        return False
    n = ctypes.c_int(0)
    return cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without a GPU skips the GPU tests instead of stopping at the first one."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _cuda_device_present():
        skip = pytest.mark.skip(reason="no CUDA device in this machine")
        for it in gpu_items:
            it.add_marker(skip)


def isotropic_C(E=1.0, nu=0.3):
    """6x6 isotropic stiffness in the reference's strain order xx,yy,zz,xy,xz,yz (DeforModelRedModule.jl:463-468)."""
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    mu = E / (2 * (1 + nu))
    C = np.zeros((6, 6))
    C[:3, :3] = lam
    C[np.arange(3), np.arange(3)] += 2 * mu
    C[3:, 3:] = mu * np.eye(3)
    return C


KAPPA3 = np.array([[1.5, 0.2, 0.1], [0.2, 2.5, 0.3], [0.1, 0.3, 3.5]])


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def fe():
    import finetools_jl_b200
    return finetools_jl_b200


@pytest.fixture(scope="session")
def gpu_ctx(fe):
    """A real device context; fails loudly (no CPU fallback) if the library or the GPU is missing."""
    return fe.GPUContext.default(0)
