"""Is the TF32 operand conversion biased? slope = <y, ref> / <ref, ref> - 1 against an fp64 reference, for the tcgen05
kernel on raw fp32 operands (hardware truncation), on operands pre-rounded to TF32 with round-to-nearest, and cuDNN."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rn_tf32(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32).view_as(t)


def main():
    from dvd_b200 import ops
    torch.backends.cudnn.allow_tf32 = True
    g = torch.Generator().manual_seed(0)
    for (N, H, W, ci, co, k) in [(4, 56, 96, 256, 256, 1), (4, 28, 48, 256, 256, 3)]:
        x = (torch.randn(N, ci, H, W, generator=g).abs() + 0.1).cuda().contiguous(memory_format=torch.channels_last)   # post-ReLU like
        w = (torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5 + 0.02).cuda()
        ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=k // 2)
        def stats(y, name):
            y = y.double()
            slope = (y * ref).sum() / (ref * ref).sum() - 1
            err = (y - ref).abs().max() / ref.abs().max()
            print('%dx%d %-28s slope %+.3e   max err / max %.3e' % (k, k, name, slope.item(), err.item()), flush=True)
        stats(ops.conv_nhwc_fwd(x, ops.pack_conv_weight(w), k), 'tcgen05 raw fp32 operands')
        stats(ops.conv_nhwc_fwd(rn_tf32(x), ops.pack_conv_weight(rn_tf32(w)), k), 'tcgen05 RN-rounded operands')
        stats(ops.conv_nhwc_fwd(x, ops.pack_conv_weight(rn_tf32(w)), k), 'tcgen05 RN weights only')
        stats(torch.nn.functional.conv2d(x, w.contiguous(memory_format=torch.channels_last), padding=k // 2), 'cuDNN TF32')
        torch.backends.cudnn.allow_tf32 = False
        stats(torch.nn.functional.conv2d(x, w.contiguous(memory_format=torch.channels_last), padding=k // 2), 'cuDNN fp32')
        torch.backends.cudnn.allow_tf32 = True


if __name__ == '__main__':
    main()
