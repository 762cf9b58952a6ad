import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "ultrasonic-communication_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def device_triples():
    return np.load(os.path.join(GOLDEN, "device_triples.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def refsim_vectors():
    return np.load(os.path.join(GOLDEN, "refsim_vectors.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def fir_taps():
    return np.load(os.path.join(GOLDEN, "fir_taps.npz"), allow_pickle=False)["taps"]
