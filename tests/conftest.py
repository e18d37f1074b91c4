import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def port_lib():
    import helpers
    return helpers.port()


@pytest.fixture(scope="session")
def ref_lib():
    import helpers
    if not helpers.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return helpers.ref()
