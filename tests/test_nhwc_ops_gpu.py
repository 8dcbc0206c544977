"""GPU: channels-last glue kernels of the depth nets (eval BatchNorm + residual + ReLU, x2 bilinear up-sampling)
against the PyTorch ops they replace, forward and backward; then the whole MiDaS mirror on the GPU (cuDNN convs +
these kernels) against the same module on the CPU."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('C,H,W', [(64, 17, 23), (256, 14, 24), (32, 8, 8), (2048, 7, 12)])
@pytest.mark.parametrize('relu,with_res', [(True, False), (True, True), (False, False)])
def test_bn_act_matches_torch(C, H, W, relu, with_res):
    from dvd_b200 import ops
    torch.manual_seed(C + H)
    bn = torch.nn.BatchNorm2d(C).cuda().eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.3); bn.running_mean.normal_(0, 0.5); bn.running_var.uniform_(0.5, 2.0)
    x = torch.randn(3, C, H, W, device='cuda').contiguous(memory_format=torch.channels_last).requires_grad_()
    res = torch.randn(3, C, H, W, device='cuda').contiguous(memory_format=torch.channels_last).requires_grad_() if with_res else None
    cot = torch.randn(3, C, H, W, device='cuda')
    y = ops.bn_act(x, bn, res, relu)
    assert y.is_contiguous(memory_format=torch.channels_last)
    (y * cot).sum().backward()
    got = [x.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone()] + ([res.grad.clone()] if with_res else [])
    x.grad = None; bn.weight.grad = None; bn.bias.grad = None
    if with_res:
        res.grad = None
    yr = bn(x) + (res if with_res else 0)
    yr = F.relu(yr) if relu else yr
    (yr * cot).sum().backward()
    ref = [x.grad, bn.weight.grad, bn.bias.grad] + ([res.grad] if with_res else [])
    assert rel_err(y, yr) < 1e-6
    for a, b in zip(got, ref):
        assert rel_err(a, b) < 2e-5


@pytest.mark.parametrize('align', [True, False])
@pytest.mark.parametrize('C,H,W', [(256, 7, 12), (128, 56, 96), (4, 2, 3), (64, 1, 5)])
def test_upsample2x_matches_torch(C, H, W, align):
    from dvd_b200 import ops
    x = torch.randn(2, C, H, W, device='cuda').contiguous(memory_format=torch.channels_last).requires_grad_()
    y = ops.upsample2x(x, align)
    cot = torch.randn_like(y)
    (y * cot).sum().backward()
    g = x.grad.clone(); x.grad = None
    yr = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=align)
    (yr * cot).sum().backward()
    assert rel_err(y, yr) < 2e-6
    assert rel_err(g, x.grad) < 2e-5
