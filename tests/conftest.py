import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "audio-to-motion-generation_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def product():
    """The product package (its directory name has hyphens, so import it by string)."""
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def pkg():
    return product()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return {n: np.load(os.path.join(d, n + "_reference.npz")) for n in ("mel", "eval", "model", "smooth", "melnfft")}
