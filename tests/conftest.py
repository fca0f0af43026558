import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (dab-radio_b200/), with libdab_b200.so built if it is stale."""
    mod = importlib.import_module("dab-radio_b200")
    mod.build()
    return mod


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def ref():
    """The reference's own sources compiled here (oracle/_ref); skip where that library does not exist."""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref/libdabref.so not built (needs /root/reference)")
    pyref.lib()
    return pyref
