import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def rel_err(a, b):
    """Tensor-normalised error max|a-b| / max|b| (point-wise relative error is
    meaningless on the near-zero elements produced by cancellation)."""
    import torch
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


@pytest.fixture(scope='session')
def reproject_golden():
    import torch
    return torch.load(os.path.join(GOLDEN, 'reproject_golden.pt'), weights_only=False)


@pytest.fixture(scope='session')
def mlp_golden():
    import torch
    return torch.load(os.path.join(GOLDEN, 'mlp_golden.pt'), weights_only=False)
