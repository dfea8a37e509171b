import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    data = np.load(path)
    return {k: torch.from_numpy(np.asarray(data[k])) for k in data.files}


@pytest.fixture(scope="session")
def det_state():
    """Deterministic weights shared with tests/golden/make_golden.py."""
    from oracle import lush_oracle as O
    return O.det_nets(O.POSTER_CFG, tag=1, sigma_gain=30.0), O.det_rbk_params(tag=7)
