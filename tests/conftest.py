import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_traj():
    return np.load(os.path.join(GOLDEN, "ekf_trajectories.npz"))


@pytest.fixture(scope="session")
def golden_edge():
    return np.load(os.path.join(GOLDEN, "edge_cases.npz"))


@pytest.fixture(scope="session")
def golden_wahba():
    return np.load(os.path.join(GOLDEN, "wahba_cases.npz"))


@pytest.fixture(scope="session")
def golden_step():
    return np.load(os.path.join(GOLDEN, "stepwise.npz"))


@pytest.fixture(scope="session")
def golden_rk4():
    return np.load(os.path.join(GOLDEN, "rk4_known_answer.npz"))


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
