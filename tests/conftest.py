import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")


@pytest.fixture(scope="session")
def built_lib():
    """Builds (if stale) and loads the product library.  Never falls back to anything else."""
    from trueno_rag_b200 import build, _lib
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def ctx(built_lib):
    from trueno_rag_b200 import api
    c = api.Context(0)
    yield c
    c.close()
