import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure every native piece exists (no-op when the .so files are current)."""
    import __graft_entry__ as g
    g.build()
    yield


@pytest.fixture(scope="session")
def H():
    from support import harness
    return harness


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; skip them cleanly if someone runs everything on a CPU box
    from support import harness
    if harness.has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
