import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def native_lib():
    """The CUDA library, built in-tree if stale (nvcc cross-compiles without a GPU)."""
    from fastace_b200 import build, lib
    if build.needs_build():
        build.build()
    return lib.load()


@pytest.fixture(scope="session")
def oracle(native_lib):
    from oracle.loader import Oracle
    return Oracle()
