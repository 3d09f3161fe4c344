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
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def bsb():
    """The product package with its CUDA library built (host-only entry points work without a GPU)."""
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "bayesspace_b200", "libbsb200.so")):
        g.build()
    import bayesspace_b200
    return bayesspace_b200
