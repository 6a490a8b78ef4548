import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "multigpu: needs at least 2 GPUs (skipped otherwise)")


@pytest.fixture(scope="session")
def built():
    """Build the product and the checker once per session (no-op when up to date)."""
    import hpc_multigpu_matrixmult_b200 as pkg
    from oracle import oracle

    pkg.build()
    oracle.build(with_ref=True)
    return pkg


@pytest.fixture(scope="session")
def capi(built):
    return built.capi


@pytest.fixture(scope="session")
def oracle(built):
    from oracle import oracle as o

    return o


@pytest.fixture(scope="session")
def gpu(capi):
    lib = capi.load()
    if lib.phpc_b200_device_count() < 1:
        pytest.fail("test marked gpu but no CUDA device is visible (there is no CPU fallback)")
    lib.phpc_b200_set_device(0)
    capi.mpi_init(0, 1)
    return lib
