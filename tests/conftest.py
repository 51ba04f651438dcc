import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (sm_100a); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def lib():
    import b200sr  # noqa: F401
    from b200sr import _lib
    return _lib
