import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): builds liboracle.so on first use."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def sla():
    """The product package; the shared library is built in-tree if needed (nvcc cross-compiles without a GPU)."""
    import sparse_linear_assignment_b200 as S
    S.build_library()
    return S
