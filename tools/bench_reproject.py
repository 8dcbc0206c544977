"""Micro-benchmark of the fused re-projection kernels (CUDA events, L2 flushed between iterations).
Algorithmic bytes/pixel: unproject fwd 16, fused fwd 32, fused bwd 48, unproject bwd 16."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, iters=20, flush=None):
    """Per-call GPU time from CUDA events. All iterations are enqueued behind a ~20 ms spin kernel before the first
    synchronisation, so the CPU (ctypes call + output allocation, ~50 us) always runs ahead of the GPU and the
    event-to-event interval contains only device work (the kernels of one call; the L2 flush sits outside it)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = []
    torch.cuda._sleep(40_000_000)
    for _ in range(iters):
        if flush is not None:
            flush_sink = flush.sum()   # read 256 MB: evicts L2 without leaving dirty lines behind
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e-3 for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main(B=64, H=224, W=384, smooth=False):
    from dvd_b200 import ops, synthetic
    dev = 'cuda'
    one = synthetic.make_batch([(4, 8)], H=H, W=W, seed=0, leading_dim=False, smooth_flow=smooth)
    rep = lambda t: t.to(dev).repeat(B, *([1] * (t.dim() - 1))).contiguous()  # noqa: E731
    flow, mask = rep(one['flow_1_2']), rep(one['mask_2'].reshape(1, H, W))
    poses = rep(ops.pack_poses_from_batch({k: v for k, v in one.items() if torch.is_tensor(v)}))
    d1, d2 = rep(synthetic.make_depths(1, H, W, seed=1)), rep(synthetic.make_depths(1, H, W, seed=2))
    sf = torch.randn(B, 3, H, W, device=dev) * 0.05
    cfg = ops.make_loss_cfg()
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)   # 256 MB, flushed by READING it (clean L2 lines)
    px = B * H * W
    res = {}
    scal = ops.reproject_loss_fwd(d1, d2, flow, mask, sf, poses, cfg)
    P = ops.unproject_fwd(d1, poses, 1)
    cases = {
        'unproject_fwd': (lambda: ops.unproject_fwd(d1, poses, 1), 16),
        'reproject_loss_fwd': (lambda: ops.reproject_loss_fwd(d1, d2, flow, mask, sf, poses, cfg), 32),
        'reproject_loss_bwd': (lambda: ops.reproject_loss_bwd(d1, d2, flow, mask, sf, poses, cfg, scal), 48),
        'unproject_bwd': (lambda: ops.unproject_bwd(P, poses, 1), 16),
    }
    for name, (fn, bpp) in cases.items():
        med, best = timeit(fn, flush=flush)
        res[name] = {'ms_median': med * 1e3, 'ms_best': best * 1e3, 'GBps_median': px * bpp / med / 1e9,
                     'GBps_best': px * bpp / best / 1e9, 'bytes_per_px': bpp}
        print(name, res[name], flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump({'B': B, 'H': H, 'W': W, 'smooth_flow': smooth, 'results': res},
              open(os.path.join(ROOT, 'gpurun_out', 'bench_reproject%s.json' % ('_smooth' if smooth else '')), 'w'), indent=1)


if __name__ == '__main__':
    main(B=int(sys.argv[1]) if len(sys.argv) > 1 else 64, smooth=len(sys.argv) > 2 and sys.argv[2] == 'smooth')
