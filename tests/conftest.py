import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(REPO, "tests", "golden", "trajectories.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def golden_hashes():
    out = {}
    with open(os.path.join(REPO, "tests", "golden", "hashes.txt")) as f:
        for line in f:
            k, v = line.rstrip("\n").rsplit(" ", 1)
            out[k] = v
    return out
