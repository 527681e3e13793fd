"""pytest configuration: markers, paths, shared helpers.

``-m "not gpu"`` : oracle vs. golden vectors / the real reference, host logic, C-ABI symbol
                  checks (no CUDA compute).
``-m gpu``       : parity tests proper -- the CUDA path (through the C-ABI) vs. the oracle.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# every output buffer is NaN-filled before a launch: unwritten nodes cannot pass by accident
os.environ.setdefault("DN_POISON_OUTPUTS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """npz accessor: g['E1.loss'] -> numpy, g.t('u') -> torch tensor."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))

    def __getitem__(self, k):
        return self.z[k]

    def t(self, k, device="cpu"):
        return torch.from_numpy(np.array(self.z[k])).to(device)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]
    return get


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).flatten().cpu()
    b = torch.as_tensor(b, dtype=torch.float64).flatten().cpu()
    return float(torch.linalg.norm(a - b) / max(float(torch.linalg.norm(b)), 1e-300))


def rel_scalar(a, b):
    a, b = float(a), float(b)
    return abs(a - b) / max(abs(b), 1e-300)
