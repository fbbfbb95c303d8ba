import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """a plain `pytest` on a box without a CUDA device skips the gpu-marked tests instead of erroring (the engine has no CPU fallback);
    `-m gpu` on such a box still fails loudly: asking for the GPU tests without a GPU is an error, not a skip"""
    if "gpu" in (config.getoption("-m") or ""):
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device (the engine has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def engine_cls():
    """the CUDA engine; loading fails loudly when the extension is missing (no CPU fallback)"""
    from ptmcmc_b200.engine import Engine
    return Engine


@pytest.fixture(scope="session")
def oracle_cls():
    from tests.oracle_binding import Oracle
    return Oracle
