#!/usr/bin/env python
"""Where do the MiDaS engine's parameter gradients differ from the fp32 reference, and is it TF32 noise or a bug?
Three gradients of the same (depth * cotangent).sum() on the GPU: (a) the functional oracle in eager PyTorch with TF32 OFF
(ground truth), (b) the same with cuDNN TF32 ON (what the reference gets on a GPU), (c) dvd_b200's engine. Per parameter:
slope - 1 (projection on the truth) and max-norm error of (b) and (c). Writes gpurun_out/r2_debug_engine.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def stats(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    den = float((b * b).sum())
    if den == 0:
        return None
    return [float((a * b).sum() / den) - 1.0, float((a - b).abs().max() / b.abs().max()), float(((a - b) ** 2).sum() ** 0.5 / den ** 0.5)]


def main():
    from dvd_b200 import synthetic
    from dvd_b200.third_party.MiDaS import MidasNet
    from oracle import depth_nets
    smooth = os.environ.get('COT', 'smooth')
    N, H, W = 2, 224, 384
    net = synthetic.seed_net_(MidasNet(non_negative=True, normalize_input=True), 0, 2000.0).eval().cuda()
    x = torch.rand(N, 3, H, W, generator=torch.Generator().manual_seed(3)).cuda()
    g = torch.Generator().manual_seed(4)
    if smooth == 'smooth':
        cot = torch.nn.functional.interpolate(torch.randn(N, 1, 8, 13, generator=g), size=(H, W), mode='bilinear').cuda() * 1e-3
    else:
        cot = torch.randn(N, 1, H, W, generator=g).cuda() * 1e-3

    def oracle_grads(tf32):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = False
        sd = {k: (v.detach().clone().requires_grad_() if (v.dtype.is_floating_point and 'running' not in k) else v)
              for k, v in net.state_dict().items()}
        d = depth_nets.midas_forward(sd, x)
        (d * cot).sum().backward()
        return d.detach(), {k: v.grad for k, v in sd.items() if getattr(v, 'grad', None) is not None}

    d0, g0 = oracle_grads(False)
    d1, g1 = oracle_grads(True)
    for p in net.parameters():
        p.grad = None
    d2 = net(x)
    (d2 * cot).sum().backward()
    g2 = {k: p.grad for k, p in net.named_parameters()}
    out = {'depth': {'cudnn_tf32': stats(d1, d0), 'engine': stats(d2.detach(), d0)}, 'params': {}}
    for k in g0:
        out['params'][k] = {'cudnn_tf32': stats(g1[k], g0[k]), 'engine': stats(g2[k].reshape(g0[k].shape), g0[k]), 'n': g0[k].numel()}
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'r2_debug_engine.json'), 'w'), indent=0)
    print('depth', out['depth'])
    keys = list(out['params'])
    pick = [k for k in keys if k.endswith('weight') and ('conv' in k or 'layer1.0' in k or 'rn' in k or 'output_conv' in k)]
    for k in pick[::4]:
        v = out['params'][k]
        print('%-55s tf32 %s   engine %s' % (k, ['%.1e' % a for a in v['cudnn_tf32']], ['%.1e' % a for a in v['engine']]))


if __name__ == '__main__':
    main()
