"""Is a halo-mode MMA slower when its A descriptor starts off an 8-row boundary? 9-tap vertical stencil (dx = 0: every tap offset is a
multiple of the 8-pixel tile row = 1024 bytes) vs 9-tap horizontal stencil (dy = 0: offsets of 1..8 rows of 128 bytes), each in halo
mode and in tap mode. Same FLOPs, same weights traffic."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dvd_b200 import conv_ops as co


def run(N, H, W, ci, co_, taps, halo, kblock=0):
    x = torch.randn(N, ci, H, W, device='cuda').contiguous(memory_format=torch.channels_last)
    w = torch.randn(len(taps), co_, kblock if kblock else ci, device='cuda')
    y = co.empty_cl(N, co_, H, W, 'cuda')
    d = co.make_desc(N, H, W, ci, H, W, co_, taps, 1, kblock, round_out=False)
    os.environ['DVD_CONV_HALO'] = '1' if halo else '0'
    for _ in range(3):
        co.conv2d_launch(d, x, w, y)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        co.conv2d_launch(d, x, w, y)
    b.record()
    torch.cuda.synchronize()
    del os.environ['DVD_CONV_HALO']
    return a.elapsed_time(b) / 10


vert = [(dy, 0, i) for i, dy in enumerate(range(-4, 5))]
horz = [(0, dx, i) for i, dx in enumerate(range(-4, 5))]
sq = [(dy, dx, (dy + 1) * 3 + dx + 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
for name, cfg in [('head 128->32 @224x384', (16, 224, 384, 128, 32, 0)), ('grouped 1024 @14x24', (16, 14, 24, 1024, 1024, 64)),
                  ('dense 256->256 @56x96', (16, 56, 96, 256, 256, 0))]:
    N, H, W, ci, co_, kb = cfg
    for tn, taps in (('vertical 9x1', vert), ('horizontal 1x9', horz), ('3x3', sq)):
        t1 = run(N, H, W, ci, co_, taps, True, kb)
        t0 = run(N, H, W, ci, co_, taps, False, kb)
        print('%-24s %-15s halo %.3f ms   taps %.3f ms' % (name, tn, t1, t0))
