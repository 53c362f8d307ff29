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
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def zlib_():
    """The product library (host-side entry points work without a GPU)."""
    from zpaqsharp_b200 import build, libzpaq
    build.build()
    return libzpaq


@pytest.fixture(scope="session")
def gpu_ctx(zlib_):
    ctx = zlib_.Context()
    yield ctx
    ctx.close()
