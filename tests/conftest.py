import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def handle():
    """A live lqrb handle on cuda:0.  No fallback: fails if the library or the GPU is missing."""
    from lqr_b200 import _lib
    h = _lib.Handle(0)
    yield h
    h.close()
