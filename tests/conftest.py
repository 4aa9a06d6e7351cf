import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gpu():
    """Binds libsparkb200.so to cuda:0.  No fallback: a missing extension or device fails the test."""
    from spark_b200 import _capi as capi
    lib = capi.init(0)
    return lib


@pytest.fixture()
def stream(gpu):
    from spark_b200.columnar import Stream
    s = Stream()
    yield s
    s.close()
