import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Build the product library and the C oracle once per session if they are stale/missing
    (no-op on the GPU box, where the prebuilt in-tree .so files travel with the snapshot)."""
    import __graft_entry__ as ge
    ge.build(quiet=True)
