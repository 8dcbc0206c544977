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


def grad_agreement(g, ref):
    """(slope - 1, relative L2 error, max-norm error) of a gradient against its reference: the projection catches a missing /
    doubled term or a wrong scale, the L2 and max-norm figures bound the noise. Reported together so that a real bug cannot
    hide behind a share-of-elements criterion."""
    import torch
    a, b = g.detach().double().cpu().reshape(-1), ref.detach().double().cpu().reshape(-1)
    den = float((b * b).sum())
    if den == 0.0:
        return 0.0, float(a.abs().max()), float(a.abs().max())
    return (float((a * b).sum() / den) - 1.0, float(((a - b) ** 2).sum().sqrt() / den ** 0.5),
            float((a - b).abs().max() / b.abs().max()))


# depth-net gradients through 101 TF32 convolution layers (random weights): the reference's own GPU path (cuDNN TF32) is 2-4 %
# (relative L2) away from the fp32 CPU gradients per tensor, slope within 1e-2 (profiles/r2_debug_engine.txt)
TF32_GRAD_SLOPE, TF32_GRAD_L2 = 0.05, 0.15
