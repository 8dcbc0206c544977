"""MiDaS inference with its dense stride-1 convolutions on the tcgen05 kernel. The integration is opt-in this round
(DVD_CONV_TC=1, see the note next to `_TC_CONV` in third_party/MiDaS.py) and so is this test: it runs only with that
variable set, last (file name) and under a hard timeout. The kernels themselves are covered unconditionally by
tests/test_conv_gpu.py."""
import os

import pytest
import torch

from conftest import rel_err

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get('DVD_CONV_TC', '0') != '1', reason='tcgen05 conv integration is opt-in: DVD_CONV_TC=1')]


@pytest.mark.timeout(180)
def test_midas_inference_on_tcgen05_convs_matches_library_path(monkeypatch):
    """MidasNet forward under no_grad runs its dense stride-1 1x1 / 3x3 convolutions on dvd_conv_nhwc_fwd with the
    BatchNorm / bias / residual / ReLU epilogues fused; with the autograd graph enabled the same module runs the
    cuDNN convolutions + glue kernels. Both are TF32: depth maps agree within 1e-3 of the tensor's max."""
    from dvd_b200 import ops, synthetic
    from dvd_b200.third_party import MiDaS
    from dvd_b200.third_party.MiDaS import MidasNet
    monkeypatch.setattr(MiDaS, '_TC_CONV', True)      # opt-in switch of this round (DVD_CONV_TC=1)
    net = MidasNet().cuda().eval()
    synthetic.seed_net_(net, 0, 2000.0)
    x = torch.rand(2, 3, 224, 384, generator=torch.Generator().manual_seed(3)).cuda()
    n0 = ops.LAUNCHES['n']
    with torch.no_grad():
        y_tc = net(x)
    assert ops.LAUNCHES['n'] - n0 > 150          # ~80 fused conv launches + glue kernels went through the C ABI
    y_lib = net(x).detach()
    assert rel_err(y_tc, y_lib) < 1e-3
