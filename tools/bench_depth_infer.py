"""MiDaS inference forward (no_grad): tcgen05 fused convolutions (default) vs the library path (DVD_CONV_TC=0 in a
second process). CUDA events over back-to-back forwards."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(nimg=8):
    from dvd_b200 import ops, synthetic
    from dvd_b200.third_party.MiDaS import MidasNet
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    net = synthetic.seed_net_(MidasNet(), 0, 2000.0).cuda().eval()
    x = torch.rand(nimg, 3, 224, 384, device='cuda')
    with torch.no_grad():
        for _ in range(3):
            net(x)
        torch.cuda.synchronize()
        n0 = ops.LAUNCHES['n']
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            net(x)
        b.record()
        torch.cuda.synchronize()
    r = {'images': nimg, 'conv_tc': os.environ.get('DVD_CONV_TC', '1'), 'fwd_ms': a.elapsed_time(b) / 10,
         'dvd_launches_per_fwd': (ops.LAUNCHES['n'] - n0) / 10}
    print(json.dumps(r), flush=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'bench_depth_infer.jsonl'), 'a') as f:
        f.write(json.dumps(r) + '\n')


if __name__ == '__main__':
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
