import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import _oracle
    return _oracle.Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference (oracle/_ref). Skips when neither the prebuilt .so nor
    /root/reference is available."""
    import _oracle
    try:
        return _oracle.Reference()
    except FileNotFoundError as e:
        pytest.skip(str(e))


@pytest.fixture(scope="session")
def phj():
    import partitionedhashjoin_b200 as p
    return p
