#!/usr/bin/env python
"""One convolution layer of the depth net in isolation (for ncu): python tools/bench_conv_layer.py KIND N H W CIN COUT K STRIDE GROUPS [iters]
KIND in fwd | dgrad | wgrad. Prints the CUDA-event time per launch and TFLOP/s."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    from dvd_b200 import conv_ops as co
    kind = sys.argv[1]
    N, H, W, ci, co_, k, stride, groups = [int(a) for a in sys.argv[2:10]]
    iters = int(sys.argv[10]) if len(sys.argv) > 10 else 10
    conv = torch.nn.Conv2d(ci, co_, k, stride=stride, padding=k // 2, groups=groups, bias=False).cuda()
    bn = torch.nn.BatchNorm2d(co_).cuda().eval()
    conv.weight.grad = torch.zeros_like(conv.weight)
    bn.weight.grad = torch.zeros_like(bn.weight)
    bn.bias.grad = torch.zeros_like(bn.bias)
    if co_ % 128 and groups == 1:
        bn = None        # swapped weight-gradient operands carry no BatchNorm
    c = co.Conv(conv, bn)
    c.pack()
    OH, OW = c.out_hw(H, W)
    x = co.round_tf32(torch.randn(N, ci, H, W, device='cuda').contiguous(memory_format=torch.channels_last))
    gy = co.round_tf32(torch.randn(N, co_, OH, OW, device='cuda').contiguous(memory_format=torch.channels_last))
    res = torch.randn(N, co_, OH, OW, device='cuda').contiguous(memory_format=torch.channels_last)
    fn = {'fwd': lambda: c.fwd(x, res=res, relu=True), 'dgrad': lambda: c.dgrad(gy, H, W, mask=x), 'wgrad': lambda: c.wgrad(x, gy, sums=True)}[kind]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) / iters
    print('%s N=%d %dx%d %d->%d k%d s%d g%d: %.1f us/launch, %.1f TFLOP/s' % (kind, N, H, W, ci, co_, k, stride, groups, t * 1e3,
                                                                               c.flops(N, OH, OW) / t / 1e9))


if __name__ == '__main__':
    main()
