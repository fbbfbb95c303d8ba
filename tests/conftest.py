import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def engine_cls():
    """the CUDA engine; loading fails loudly when the extension is missing (no CPU fallback)"""
    from ptmcmc_b200.engine import Engine
    return Engine


@pytest.fixture(scope="session")
def oracle_cls():
    from tests.oracle_binding import Oracle
    return Oracle
