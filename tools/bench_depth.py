"""Depth-net (MiDaS, cuDNN path) forward+backward variants: memory format / CUDA-graph capture."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ev(fn, iters=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main(nimg=8):
    from dvd_b200 import synthetic
    from dvd_b200.third_party.MiDaS import MidasNet
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    res = {}
    x = torch.rand(nimg, 3, 224, 384, device='cuda')
    for fmt_name, fmt in (('nchw', torch.contiguous_format), ('channels_last', torch.channels_last)):
        net = synthetic.seed_net_(MidasNet(non_negative=True, normalize_input=True), 0, 2000.0).cuda().eval()
        if fmt_name == 'channels_last':
            net = net.to(memory_format=torch.channels_last)
        xi = x.contiguous(memory_format=fmt)

        def fb():
            for p in net.parameters():
                p.grad = None
            d = net(xi)
            d.sum().backward()
        res[fmt_name + '_eager_ms'] = ev(fb)
        with torch.no_grad():
            res[fmt_name + '_fwd_ms'] = ev(lambda: net(xi))
        try:
            g = torch.cuda.make_graphed_callables(net, (xi.clone().requires_grad_(False),))

            def fbg():
                for p in net.parameters():
                    p.grad = None
                d = g(xi)
                d.sum().backward()
            res[fmt_name + '_graphed_ms'] = ev(fbg)
        except Exception as e:  # noqa: BLE001
            res[fmt_name + '_graphed_ms'] = 'failed: %s' % str(e)[:200]
        print(fmt_name, {k: v for k, v in res.items() if k.startswith(fmt_name)}, flush=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'bench_depth.json'), 'w'), indent=1)


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
