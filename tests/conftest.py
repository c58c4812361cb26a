import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_sessionstart(session):
    """Make sure the in-tree shared library exists (nvcc cross-compiles without a GPU; a no-op when it is
    up to date) so that the C-ABI export test and the GPU tests see the current sources."""
    try:
        import __graft_entry__ as ge
        ge.build()
    except Exception as e:  # noqa: BLE001 - report, let the tests that need the library fail loudly
        print(f"[conftest] build() failed: {e}")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
