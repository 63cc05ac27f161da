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
    from tests import helpers
    return helpers.Oracle()


@pytest.fixture(scope="session")
def hostsim():
    from tests import helpers
    return helpers.HostSim()


@pytest.fixture(scope="session")
def gpu_ctx():
    import zstandard_b200 as zb
    ctx = zb.Context(max_batch_bytes=64 << 20)
    yield ctx
    ctx.close()
