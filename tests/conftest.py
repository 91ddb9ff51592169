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


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not failed) when the process has no CUDA device, whatever fixture they use."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in gpu_items:
        it.add_marker(skip)


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
