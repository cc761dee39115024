import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def handle():
    """One libgpb200 handle for the whole GPU session (fails loudly without a device)."""
    from gptest_b200 import _lib
    h = _lib.Handle(0)
    yield h
    h.close()
