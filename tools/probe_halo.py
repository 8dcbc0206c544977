"""Does a tcgen05 SWIZZLE_128B K-major descriptor tolerate a start address that is shifted by whole 128-byte rows (not a multiple
of 8) and 8-row groups that are more than 1024 bytes apart? (halo-resident convolution tiles: one TMA box per 32-channel chunk,
the taps of a kxk stencil become descriptor offsets). Prints max error per (shift, pitch, base-offset mode)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dvd_b200 import _lib

def tf32(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)

lib = _lib.load()
P = ctypes.c_void_p
R, N = 400, 64
g = torch.Generator().manual_seed(0)
A = tf32(torch.randn(R, 32, generator=g)).cuda()
B = tf32(torch.randn(N, 32, generator=g)).cuda()
for sbo in (8, 10, 18):
    for shift in (0, 1, 3, 8, 10, 11, 21):
        for bo in (0, 1):
            D = torch.zeros(128, N, device='cuda')
            _lib.check(lib.dvd_selftest_halo(P(A.data_ptr()), P(B.data_ptr()), P(D.data_ptr()), R, N, shift, sbo, bo, None), 'halo')
            torch.cuda.synchronize()
            rows = torch.tensor([shift + (m // 8) * sbo + m % 8 for m in range(128)], device='cuda')
            ref = A[rows].double() @ B.double().t()
            err = float((D.double() - ref).abs().max() / ref.abs().max())
            print('sbo_rows %2d shift %2d base_offset_mode %d  err %.2e  %s' % (sbo, shift, bo, err, 'OK' if err < 1e-5 else 'WRONG'))
