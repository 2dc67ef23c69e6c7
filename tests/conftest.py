import os
import sys
import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 (B200) device")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def predict_weights():
    from sentinel_tree_cover_b200.weights import load_npz
    return load_npz(os.path.join(GOLDEN, "weights_predict_172.npz"))


@pytest.fixture(scope="session")
def sr_weights():
    from sentinel_tree_cover_b200.weights import load_npz
    return load_npz(os.path.join(GOLDEN, "weights_superresolve.npz"))


@pytest.fixture(scope="session")
def sess(predict_weights, sr_weights):
    """One StcSession per test run (GPU only).  Fails loudly without libstc.so / a B200."""
    from sentinel_tree_cover_b200.api import StcSession
    s = StcSession(0, predict_weights=predict_weights, superresolve_weights=sr_weights)
    yield s
    s.close()
