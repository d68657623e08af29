import os
import sys
import importlib
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "gpu_unverified: CUDA code compiled but not yet run on hardware; skips itself without a device")


def pytest_collection_modifyitems(config, items):
    """B200_RUN_UNVERIFIED=1 puts the hardware-pending tests (marker gpu_unverified) under `-m gpu` as well, so that one
    `pytest -m gpu` run on a GPU box covers them before their markers are switched in the source."""
    if os.environ.get("B200_RUN_UNVERIFIED") == "1":
        for it in items:
            if it.get_closest_marker("gpu_unverified"):
                it.add_marker(pytest.mark.gpu)


@pytest.fixture(scope="session")
def b200():
    """The harness package (directory name has a hyphen, so import by string)."""
    return importlib.import_module("mp-gadget_b200")


@pytest.fixture(scope="session")
def ics():
    return importlib.import_module("mp-gadget_b200.ics")


@pytest.fixture(scope="session")
def engine(b200):
    e = b200.Engine(0)
    yield e
    e.close()
