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


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """The CUDA library is compiled in-tree (nvcc cross-compiles without a GPU); tests never run against a
    missing or stale build silently."""
    from glomeruli_segmentation_b200 import _lib
    src_dir = os.path.join(ROOT, "glomeruli_segmentation_b200", "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh", ".sh"))]
    srcs.append(os.path.join(ROOT, "include", "espnet_b200.h"))
    if (not os.path.isfile(_lib.LIB_PATH)) or os.path.getmtime(_lib.LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        _lib.build()
    return _lib.LIB_PATH
