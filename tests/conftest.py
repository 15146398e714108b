import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        from seabreeze_param_b200 import _capi
        return _capi.load().sbd_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: there is no CPU fallback to test.
    pass


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.set_math(1)
    O.set_threads(1)
    return O
