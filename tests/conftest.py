import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure).  Prefers the build that carries the reference's own detector."""
    from tests import oracle_binding
    return oracle_binding.load()


@pytest.fixture(scope="session")
def mfk():
    import mofreak_b200
    mofreak_b200.lib()
    return mofreak_b200
