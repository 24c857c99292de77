import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")


@pytest.fixture(scope="session")
def gtop_lib():
    """libgtop_b200.so must be present and must see a GPU for the -m gpu tests: fail, do not skip."""
    from grad_traj_optimization_b200 import _capi
    L = _capi.lib()
    assert L.gtop_device_count() >= 1, "no CUDA device visible: the gpu-marked tests need a B200"
    return L


@pytest.fixture(scope="session")
def gtop_lib_cpu():
    """The product library for its host-only entry points (no device needed, no compute kernels launched)."""
    from grad_traj_optimization_b200 import _capi
    return _capi.lib()
