import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine bound to cuda:0.  No fallback: raises without a GPU."""
    import decaf377_b200 as d
    d.init(0)
    yield d
