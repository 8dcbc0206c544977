"""tcgen05 TF32 convolution (dvd_conv_nhwc_fwd) vs cuDNN (torch, channels_last, TF32 allowed) on the MiDaS shapes that
the kernel covers. CUDA events, inputs of all iterations enqueued behind a spin kernel (no host time in the interval)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(20_000_000)
    evs = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e-3 for a, b in evs)
    return ts[len(ts) // 2]


def main(images=16):
    from dvd_b200 import ops
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    shapes = [  # H, W, Cin, Cout, k   (384x224 input)
        (56, 96, 256, 256, 1), (56, 96, 64, 256, 1), (28, 48, 512, 512, 1), (14, 24, 1024, 1024, 1),
        (7, 12, 2048, 2048, 1), (56, 96, 256, 256, 3), (28, 48, 256, 256, 3), (14, 24, 256, 256, 3),
        (112, 192, 256, 128, 3),
    ]
    res = []
    for H, W, ci, co, k in shapes:
        x = torch.randn(images, ci, H, W, device='cuda').contiguous(memory_format=torch.channels_last)
        w = (torch.randn(co, ci, k, k, device='cuda') / (ci * k * k) ** 0.5).contiguous(memory_format=torch.channels_last)
        wp = ops.pack_conv_weight(w)
        flops = 2.0 * images * H * W * ci * co * k * k
        t_mine = timeit(lambda: ops.conv_nhwc_fwd(x, wp, k))
        t_lib = timeit(lambda: torch.nn.functional.conv2d(x, w, padding=k // 2))
        r = {'H': H, 'W': W, 'Cin': ci, 'Cout': co, 'k': k, 'images': images, 'tcgen05_us': t_mine * 1e6,
             'cudnn_us': t_lib * 1e6, 'tcgen05_TFLOPs': flops / t_mine / 1e12, 'cudnn_TFLOPs': flops / t_lib / 1e12,
             'io_GBps': 4.0 * images * H * W * (ci + co) / t_mine / 1e9}
        res.append(r)
        print(r, flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'bench_conv.json'), 'w'), indent=1)


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16)
