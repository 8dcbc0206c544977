"""replays the forward cases of tests/test_conv2d_gpu.py in order; reports flags + error structure per case"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.nn.functional as F
from dvd_b200 import conv_ops as co
import test_conv2d_gpu as T
from test_conv2d_gpu import make_conv, make_bn, tf32, cl, gen

def flags():
    ws = co._WORKSPACE.get(0)
    if ws is None: return None
    torch.cuda.synchronize()
    f = ws[:256].view(torch.int32)
    return [int(i) for i in torch.nonzero(f).flatten().tolist()]

for idx, case in enumerate(T.FWD_CASES):
    N, H, W, ci, co_, k, stride, groups, use_bn, bias, has_res, has_res2, relu, has_mask = case
    seed = 1000 * k + ci + co_ + H + 7 * stride + groups
    g = gen(seed)
    conv = make_conv(ci, co_, k, stride, groups, bias, seed + 1)
    bn = make_bn(co_, seed + 2) if use_bn else None
    x = tf32(torch.randn(N, ci, H, W, generator=g))
    OH, OW = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
    res = torch.randn(N, co_, OH, OW, generator=g) if has_res else None
    res2 = torch.randn(N, co_, OH, OW, generator=g) if has_res2 else None
    mask = torch.randn(N, co_, OH, OW, generator=g) if has_mask else None
    conv, bn = conv.cuda(), (bn.cuda() if bn is not None else None)
    c = co.Conv(conv, bn)
    c.pack(need_bwd=False)
    d = co.make_desc(N, H, W, ci, OH, OW, co_, co.fwd_taps(k, k // 2), stride, c.kblock, relu=relu, round_out=False,
                     bn_eps=bn.eps if bn is not None else 0.0)
    args = (d, cl(x), c.w_fwd)
    kw = dict(bias=conv.bias, bn=c._bn_fwd(), res=cl(res) if has_res else None, res2=cl(res2) if has_res2 else None,
              mask=cl(mask) if has_mask else None)
    f0 = flags()
    y1 = co.conv2d_launch(*args, co.empty_cl(N, co_, OH, OW, 'cuda'), **kw)
    f1 = flags()
    os.environ['DVD_CONV_STREAMK'] = '0'
    y0 = co.conv2d_launch(*args, co.empty_cl(N, co_, OH, OW, 'cuda'), **kw)
    del os.environ['DVD_CONV_STREAMK']
    torch.cuda.synchronize()
    err = (y1 - y0).abs()
    mx = float(y0.abs().max())
    bad = err > 1e-3 * mx
    print(idx, case[:8], 'flags before', f0, 'after', f1, 'err', float(err.max()) / mx, 'badfrac', float(bad.float().mean()))
    if bad.any():
        e = bad.permute(0, 2, 3, 1).float()
        print('   per image', e.mean(dim=(1, 2, 3)).cpu().numpy().round(3))
        print('   per h', e.mean(dim=(0, 2, 3)).cpu().numpy().round(2))
        print('   per w', e.mean(dim=(0, 1, 3)).cpu().numpy().round(2))
        print('   per 32ch', e.reshape(-1, co_ // 32 if co_ >= 32 else 1, min(32, co_)).mean(dim=(0, 2)).cpu().numpy().round(2))
