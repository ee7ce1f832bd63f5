import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "espnet_golden.npz"))


def load_fold_state_dict(fold):
    z = np.load(os.path.join(GOLDEN, "weights_fold%d.npz" % fold))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def fold_sd():
    cache = {}

    def get(fold):
        if fold not in cache:
            cache[fold] = load_fold_state_dict(fold)
        return cache[fold]
    return get
