import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure libgvdb.so and the oracle exist (cross-compiles on a CPU box)."""
    import __graft_entry__ as g
    from grape_vector_db_b200 import _ffi
    if not os.path.exists(_ffi.LIB_PATH):
        g.build()
    return True
