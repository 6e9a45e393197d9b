import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_pkg():
    return importlib.import_module("senweaver-ide_b200")


@pytest.fixture(scope="session")
def apo():
    return load_pkg()


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def engine(apo):
    """A live engine on cuda:0 — only -m gpu tests request it; it raises (never falls back)
    when the CUDA library or the device is missing."""
    e = apo.Engine(0)
    yield e
    e.close()
