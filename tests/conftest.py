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


def pytest_terminal_summary(terminalreporter):
    """With D377_DEBUG_REPORT=1 (set by the debug-build test) report how many points the
    on-curve predicate compiled into the kernels has checked, and how many failed."""
    if os.environ.get("D377_DEBUG_REPORT", "0") in ("", "0"):
        return
    try:
        import decaf377_b200 as d
        d.init(0)
        failures, checked = d.debug_counts()
        terminalreporter.write_line("D377_DEBUG build=%d checked=%d failures=%d"
                                    % (d.debug_build(), checked, failures))
    except Exception as ex:  # noqa: BLE001
        terminalreporter.write_line("D377_DEBUG unavailable: %r" % (ex,))
