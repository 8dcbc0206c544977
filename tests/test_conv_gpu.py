"""GPU parity of the tcgen05 TF32 convolution (dvd_conv_nhwc_fwd, through the C ABI) against torch's convolution in
fp64 on the same inputs. Both operands are rounded to TF32 with round-to-nearest (weights when they are packed, the
activation tile in shared memory before the MMAs) - what cuDNN gives the reference on a GPU by default - and
accumulated in fp32: tolerance 1e-3 of the tensor's max, and no multiplicative bias (the tensor core alone would
truncate the operands: slope -7e-4 per layer, tools/debug_conv_bias.py)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

CASES = [
    # N, H,  W,  Cin, Cout, k, affine, res, relu
    (2, 28, 48, 256, 512, 1, True, True, True),      # ResNeXt conv3 + bn3 + identity + relu
    (1, 56, 96, 64, 256, 1, True, False, False),     # downsample branch
    (3, 7, 12, 2048, 256, 1, False, False, False),   # ragged pixel count (252), K = 2048
    (2, 56, 96, 256, 256, 3, False, False, False),   # layer1_rn / RCU conv, tile 4 x 32
    (1, 28, 48, 256, 256, 3, True, True, True),      # tile 8 x 16, H not a multiple of 8
    (2, 14, 24, 512, 256, 3, True, False, True),     # tile 16 x 8, ragged H
    (1, 112, 192, 256, 128, 3, True, False, False),  # head conv, Cout = 128
    (1, 30, 50, 128, 32, 3, True, False, True),      # W not a multiple of any tile width, Cout = 32
    (1, 9, 13, 32, 16, 3, False, True, False),       # tiny, Cout = 16
]


@pytest.mark.parametrize('case', CASES)
def test_conv_matches_torch_fp64(case):
    from dvd_b200 import ops
    N, H, W, ci, co, k, affine, has_res, relu = case
    g = torch.Generator().manual_seed(1000 * k + ci + co + H)
    x = torch.randn(N, ci, H, W, generator=g)
    w = torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5
    bias = torch.randn(co, generator=g) if (affine and k == 3) else None          # decoder convs carry a bias
    bn = None
    if affine and not (k == 3 and H % 4 == 0):                                   # ... the encoder ones a BatchNorm
        bn = (torch.rand(co, generator=g) + 0.5, torch.randn(co, generator=g), torch.randn(co, generator=g) * 0.1,
              torch.rand(co, generator=g) + 0.5, 1e-5)
    res = torch.randn(N, co, H, W, generator=g) if has_res else None
    ref = torch.nn.functional.conv2d(x.double(), w.double(), bias.double() if bias is not None else None, padding=k // 2)
    if bn is not None:
        ref = torch.nn.functional.batch_norm(ref, bn[2].double(), bn[3].double(), bn[0].double(), bn[1].double(), False, 0.0, bn[4])
    if has_res:
        ref = ref + res.double()
    if relu:
        ref = ref.relu()
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    y = ops.conv_nhwc_fwd(xc, ops.pack_conv_weight(w.cuda()), k, bias.cuda() if bias is not None else None,
                          tuple(t.cuda() for t in bn[:4]) + (bn[4],) if bn is not None else None,
                          res.cuda() if has_res else None, relu)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert rel_err(y, ref) < 1e-3
    if not relu and not has_res and not affine:
        slope = ((y.double().cpu() * ref).sum() / (ref * ref).sum() - 1).item()
        assert abs(slope) < 1e-4, slope


def test_conv_dgrad_image_is_the_adjoint():
    """<conv(x, w), g> == <x, conv(g, w_dgrad)>: the data gradient runs through the same kernel."""
    from dvd_b200 import ops
    g0 = torch.Generator().manual_seed(5)
    N, H, W, ci, co = 2, 28, 48, 64, 128
    x = torch.randn(N, ci, H, W, generator=g0).cuda().contiguous(memory_format=torch.channels_last)
    gy = torch.randn(N, co, H, W, generator=g0).cuda().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(co, ci, 3, 3, generator=g0) / (ci * 9) ** 0.5).cuda()
    y = ops.conv_nhwc_fwd(x, ops.pack_conv_weight(w), 3)
    gx = ops.conv_nhwc_fwd(gy, ops.pack_conv_weight(w, dgrad=True), 3)
    gx_ref = torch.nn.grad.conv2d_input(x.shape, w.double(), gy.double(), padding=1)
    assert rel_err(gx, gx_ref) < 1e-3
    a, b = (y.double() * gy.double()).sum().item(), (x.double() * gx.double()).sum().item()
    assert abs(a - b) <= 2e-3 * max(abs(a), abs(b), 1.0)


def test_unsupported_shapes_fail_loudly():
    from dvd_b200 import ops
    x = torch.randn(1, 3, 16, 16).cuda().contiguous(memory_format=torch.channels_last)
    w = torch.randn(32, 3, 3, 3).cuda()
    with pytest.raises(RuntimeError):
        ops.conv_nhwc_fwd(x, ops.pack_conv_weight(w), 3)


WGRAD_CASES = [
    # N, H,  W,  Cin,  Cout, k, channels_last weight gradient
    (2, 28, 48, 64, 128, 3, False),
    (1, 56, 96, 256, 256, 3, True),      # 16-byte vector reductions (Cin stride 1)
    (3, 7, 12, 1024, 256, 1, True),      # 1x1: K = 252 pixels (ragged), four in-channel tiles
    (1, 30, 50, 32, 128, 3, False),      # ragged spatial tiles, N = 32
    (2, 14, 24, 512, 128, 1, False),
]


@pytest.mark.parametrize('case', WGRAD_CASES)
def test_conv_wgrad_matches_torch_fp64(case):
    """dvd_conv_nhwc_wgrad (both operands MN-major TF32, split-K, fp32 reductions) vs torch.nn.grad.conv2d_weight in
    fp64. The operands are activations and are not re-rounded, so the tensor core's TF32 truncation leaves a uniform
    -7e-4 scale on the result (csrc/conv_tc.cu): tolerance 2e-3; the call accumulates into its destination."""
    from dvd_b200 import ops
    N, H, W, ci, co, k, cl = case
    g = torch.Generator().manual_seed(77 + ci + co + H)
    x = torch.randn(N, ci, H, W, generator=g)
    gy = torch.randn(N, co, H, W, generator=g)
    ref = torch.nn.grad.conv2d_weight(x.double(), (co, ci, k, k), gy.double(), padding=k // 2)
    dw = torch.zeros(co, ci, k, k, device='cuda')
    if cl:
        dw = dw.contiguous(memory_format=torch.channels_last)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    gc = gy.cuda().contiguous(memory_format=torch.channels_last)
    ops.conv_nhwc_wgrad(xc, gc, dw)
    assert rel_err(dw, ref) < 2e-3
    ops.conv_nhwc_wgrad(xc, gc, dw)          # accumulates
    assert rel_err(dw, 2 * ref) < 2e-3
