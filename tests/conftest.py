import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    import dcgan_super_resolution_b200 as dsr
    c = dsr.Context(device=0, precision="strict")
    yield c
    c.close()


@pytest.fixture(scope="session")
def ctx_fast():
    import dcgan_super_resolution_b200 as dsr
    c = dsr.Context(device=0, precision="tf32")
    yield c
    c.close()
