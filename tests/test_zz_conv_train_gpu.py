"""ops.ConvTc (forward, data gradient, weight gradient on the tcgen05 kernels behind one autograd Function) and its use in
the MiDaS mirror. STAGED FOR ROUND 2 and opt-in like the code it covers (DVD_CONV_TC_TRAIN=1): written after the round's
GPU budget was spent, never executed on a GPU so far. The three kernels underneath are covered unconditionally by
tests/test_conv_gpu.py."""
import os

import pytest
import torch

from conftest import rel_err

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get('DVD_CONV_TC_TRAIN', '0') != '1',
                                 reason='staged for round 2: DVD_CONV_TC_TRAIN=1')]


@pytest.mark.timeout(180)
@pytest.mark.parametrize('case', [(2, 28, 48, 64, 128, 3, True), (1, 56, 96, 256, 256, 1, False)])
def test_convtc_autograd_matches_torch_fp64(case):
    from dvd_b200 import ops
    N, H, W, ci, co, k, has_bias = case
    g = torch.Generator().manual_seed(9)
    x = torch.randn(N, ci, H, W, generator=g)
    w = torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5
    b = torch.randn(co, generator=g) if has_bias else None
    gy = torch.randn(N, co, H, W, generator=g)
    xr, wr = x.double().requires_grad_(), w.double().requires_grad_()
    br = b.double().requires_grad_() if has_bias else None
    yr = torch.nn.functional.conv2d(xr, wr, br, padding=k // 2)
    yr.backward(gy.double())
    xg = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    wg = w.cuda().requires_grad_()
    bg = b.cuda().requires_grad_() if has_bias else None
    y = ops.ConvTc.apply(xg, wg, bg)
    y.backward(gy.cuda())
    assert rel_err(y, yr) < 1e-3
    assert rel_err(xg.grad, xr.grad) < 1e-3
    assert rel_err(wg.grad, wr.grad) < 2e-3
    if has_bias:
        assert rel_err(bg.grad, br.grad) < 1e-4


@pytest.mark.timeout(300)
def test_midas_training_gradients_on_convtc_match_library_path(monkeypatch):
    from dvd_b200 import synthetic
    from dvd_b200.third_party import MiDaS
    from dvd_b200.third_party.MiDaS import MidasNet
    net = MidasNet().cuda().eval()
    synthetic.seed_net_(net, 0, 2000.0)
    x = torch.rand(1, 3, 224, 384, generator=torch.Generator().manual_seed(3)).cuda()
    names = ['pretrained.layer1.4.0.conv1.weight', 'pretrained.layer3.5.conv3.weight', 'scratch.refinenet2.resConfUnit1.conv1.weight']
    params = dict(net.named_parameters())

    def grads(flag):
        monkeypatch.setattr(MiDaS, '_TC_CONV_TRAIN', flag)
        net.zero_grad(set_to_none=True)
        d = net(x)
        (1.0 / d).mean().backward()
        return d.detach(), {n: params[n].grad.clone() for n in names}
    d_lib, g_lib = grads(False)
    d_tc, g_tc = grads(True)
    assert rel_err(d_tc, d_lib) < 1e-3
    for n in names:
        assert rel_err(g_tc[n], g_lib[n]) < 5e-3, n
