import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _pkg(name):
    return importlib.import_module("3dline-slam_b200." + name)


@pytest.fixture(scope="session")
def scene_mod():
    return _pkg("scene")


@pytest.fixture(scope="session")
def api():
    return _pkg("api")


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    oracle_py.build()
    return oracle_py


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
