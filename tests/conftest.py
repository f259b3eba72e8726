import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle.lib()


@pytest.fixture(scope="session")
def hmref():
    """The reference's own compiled TLibCommon (oracle/_ref/libhmref.so).  Absent on boxes
    where it was never built; tests that need it skip and the committed goldens take over."""
    import oracle
    oracle.build()
    r = oracle.ref()
    if r is None:
        pytest.skip("oracle/_ref/libhmref.so not built (needs /root/reference at build time)")
    return r
