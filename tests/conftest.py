import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


@pytest.fixture(scope="session")
def lipnet_sd():
    from oracle import lipnet_ref
    return lipnet_ref.init_lipnet_state(39, 256, seed=0)


@pytest.fixture(scope="session")
def det_sd():
    from oracle import sweep_ref
    return sweep_ref.init_detector_state(13864, 512, seed=1)
