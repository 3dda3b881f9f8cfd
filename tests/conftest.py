import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ml100k():
    import mrs_b200  # noqa: F401
    from mrs_b200 import synth
    return synth.cached("ml100k")


@pytest.fixture(scope="session")
def small():
    import mrs_b200  # noqa: F401
    from mrs_b200 import synth
    return synth.small()
