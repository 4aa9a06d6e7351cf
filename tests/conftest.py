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


@pytest.fixture()
def sbconfig(gpu):
    """sb_config_set with automatic restore (replaces the environment knobs of round 1)."""
    from spark_b200 import _capi as capi
    saved = {}

    def set_(key, value):
        if key not in saved:
            saved[key] = capi.config_get(key)
        capi.config_set(key, value)
    yield set_
    for k, v in saved.items():
        capi.config_set(k, v)
