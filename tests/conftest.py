import importlib
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(autouse=True)
def _seed_python_random():
    """The reference draws its random starts from Python's ``random`` (environment.py:35-40) and
    so does the mirror: seed it per test so that every run of the suite sees the same episodes."""
    random.seed(20261018)


@pytest.fixture(scope="session")
def pkg():
    """The product package (its directory name has a hyphen)."""
    return importlib.import_module("async-rl-tensorflow_b200")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "preprocess_golden.npz"))


@pytest.fixture(scope="session")
def cuda(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    pkg._cabi.init("cuda:0")
    return torch.device("cuda:0")
