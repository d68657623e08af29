import os
import sys
import importlib
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
