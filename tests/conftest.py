import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure the in-tree libraries exist (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as G
    G.build(force=False)
    return True


@pytest.fixture(scope="session")
def B(built):
    import bsw_b200
    return bsw_b200


@pytest.fixture(scope="session")
def O(built):
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def ctx(B):
    c = B.Context()
    yield c
    c.close()
