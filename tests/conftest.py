import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure libryolo.so / the oracle exist (no-op when __graft_entry__.build() already ran)."""
    lib = os.path.join(REPO, "rotate-yolov3_b200", "libryolo.so")
    orc = os.path.join(REPO, "oracle", "librbox_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import __graft_entry__
        __graft_entry__.build()
    yield
