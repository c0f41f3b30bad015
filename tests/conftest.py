import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle.bindings import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (Kokkos::OpenMP, 4 threads); skipped where oracle/_ref was not built."""
    from oracle.bindings import Ref, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref/libkokkos_ref_omp.so not built on this machine")
    return Ref(4)


@pytest.fixture(scope="session")
def space():
    """One B200 execution-space instance for the GPU tests; fails loudly if the extension is missing."""
    import kokkos_b200 as kb
    s = kb.B200(0)
    yield s
    s.finalize()
