import importlib
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


def load_pkg():
    """The package directory is `ego-moment-cle-vit_b200` (not an identifier): import by name."""
    return importlib.import_module("ego-moment-cle-vit_b200")


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def params_of(rec):
    return {k[2:]: v for k, v in rec.items() if k.startswith("p:")}


def make_inputs(B, N, D, seed=1234):
    """Same synthetic token generator as tests/golden/make_golden.py."""
    import torch
    g = torch.Generator().manual_seed(seed)
    anchor = torch.randn(B, N, D, generator=g)
    positive = anchor + 0.5 * torch.randn(B, N, D, generator=g)
    return anchor, positive


def make_structured_inputs(B, N, D, seed=1234, rank=16):
    """Same "structured" generator as tests/golden/make_golden.py."""
    import torch
    g = torch.Generator().manual_seed(seed)
    s = 0.5 + 1.5 * torch.rand(B, 1, 1, generator=g)
    m = torch.randn(B, 1, D, generator=g)
    L = torch.randn(B, N, rank, generator=g)
    R = torch.randn(B, D, rank, generator=g)
    anchor = s * (m + torch.bmm(L, R.transpose(1, 2)) + 0.3 * torch.randn(B, N, D, generator=g))
    positive = anchor + 0.5 * s * torch.randn(B, N, D, generator=g)
    return anchor, positive


def rel_err(x, ref):
    x = np.asarray(x, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.linalg.norm(x - ref) / (np.linalg.norm(ref) + 1e-300))
