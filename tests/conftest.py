import os
import sys
import importlib.util

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")


def _load_pkg():
    import fxb200  # root shim: loads featureextraction.jl_b200/ as module "fxb200"
    return fxb200


@pytest.fixture(scope="session")
def fx():
    return _load_pkg()


@pytest.fixture(scope="session")
def gpu_ctx(fx):
    """One library context on cuda:0 for the whole GPU session."""
    ctx = fx.Context(device=0)
    yield ctx
    ctx.close()
