"""GPU: tcgen05 building blocks (UMMA descriptors, SS/TS operand modes, bf16x3 split) vs torch fp32."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def run_selftest(K, N, mode, passes, seed=0):
    from dvd_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(128, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    D = torch.zeros(128, N, device='cuda')
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.dvd_selftest_umma(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()),
                                     ctypes.c_void_p(D.data_ptr()), K, N, mode, passes, st), 'dvd_selftest_umma')
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    return float((D.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize('mode', [0, 1, 2, 3])
@pytest.mark.parametrize('K,N', [(64, 256), (128, 256), (128, 16), (64, 64), (64, 144)])
def test_umma_bf16x3_is_fp32_grade(K, N, mode):
    assert run_selftest(K, N, mode, 3) < 2e-5


@pytest.mark.parametrize('mode', [0, 1, 2])
def test_umma_single_pass_is_bf16_grade(mode):
    e = run_selftest(128, 256, mode, 1)
    assert 1e-4 < e < 2e-2
