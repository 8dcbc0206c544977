"""Kernel-time table (torch profiler, CUDA activities) of the cuDNN depth net fwd+bwd and of a full step."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(nimg=8, cl=False):
    from dvd_b200 import synthetic
    from dvd_b200.third_party.MiDaS import MidasNet
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    net = synthetic.seed_net_(MidasNet(non_negative=True, normalize_input=True), 0, 2000.0).cuda().eval()
    x = torch.rand(nimg, 3, 224, 384, device='cuda')
    if cl:
        net = net.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)

    def fb():
        for p in net.parameters():
            p.grad = None
        net(x).sum().backward()
    for _ in range(3):
        fb()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fb()
        torch.cuda.synchronize()
    rows = prof.key_averages()
    tot = sum(r.device_time_total for r in rows)
    print('total device us', tot, 'kernels', sum(r.count for r in rows))
    for r in sorted(rows, key=lambda r: -r.device_time_total)[:32]:
        print('%9.0f us %5.1f%% %5d  %s' % (r.device_time_total, 100 * r.device_time_total / tot, r.count, r.key[:110]))


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8, cl=len(sys.argv) > 2 and sys.argv[2] == 'cl')
