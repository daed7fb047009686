import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import hrmo

    hrmo.lib()
    return hrmo


@pytest.fixture(scope="session")
def gpu_ctx():
    """A live hrmgpu context on cuda:0.  Fails (does not skip) when the CUDA library or device is missing."""
    from hrm_b200 import capi

    ctx = capi.Context(0)
    yield ctx
    ctx.close()
