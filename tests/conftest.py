import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# `-m gpu` on a box without a GPU fails loudly (SBD_ECUDA from every entry point): there is no CPU fallback to test,
# so nothing is skipped here.


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.set_math(1)
    O.set_threads(1)
    return O
