import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One libiak3d context on cuda:0 for the GPU tests (fails loudly without a GPU: no fallback)."""
    import iak3d_b200
    c = iak3d_b200.Context(0)
    yield c
    c.close()
