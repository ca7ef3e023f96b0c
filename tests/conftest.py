import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def b2():
    """The host package (its directory name has a hyphen, hence importlib)."""
    return importlib.import_module("zip-ada_b200")


@pytest.fixture(scope="session")
def b2mod():
    return b2()


@pytest.fixture(scope="session")
def enc9(b2mod):
    e = b2mod.Encoder(b2mod.block_900k, 0)
    yield e
    e.close()
