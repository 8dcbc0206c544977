"""stream-K debugging: same launch with DVD_CONV_STREAMK on / off, error broken down by 128-pixel tile and 32-channel block"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from dvd_b200 import conv_ops as co
from test_conv2d_gpu import make_conv, tf32, cl, gen


def run(N, H, W, ci, co_, k, bias, mask, relu, reps=1):
    g = gen(5)
    conv = make_conv(ci, co_, k, 1, 1, bias, 6).cuda()
    c = co.Conv(conv, None)
    c.pack(need_bwd=False)
    x = cl(tf32(torch.randn(N, ci, H, W, generator=g)))
    m = cl(torch.randn(N, co_, H, W, generator=g)) if mask else None
    d = co.make_desc(N, H, W, ci, H, W, co_, co.fwd_taps(k, k // 2), 1, 0, relu=relu, round_out=False)
    outs = {}
    for mode in ('0', '1', '1', '0'):
        os.environ['DVD_CONV_STREAMK'] = mode
        for _ in range(reps):
            y = co.conv2d_launch(d, x, c.w_fwd, co.empty_cl(N, co_, H, W, 'cuda'), conv.bias, None, None, None, m)
        torch.cuda.synchronize()
        outs.setdefault(mode, []).append(y.clone())
    a, b = outs['0'][0], outs['1'][0]
    err = (a - b).abs()
    print('case', (N, H, W, ci, co_, k, bias, mask, relu), 'max', float(err.max()), 'ref max', float(a.abs().max()),
          'sk repeat equal', bool(torch.equal(outs['1'][0], outs['1'][1])), 'dp repeat equal', bool(torch.equal(outs['0'][0], outs['0'][1])))
    if float(err.max()) > 1e-3 * float(a.abs().max()):
        e = err.permute(0, 2, 3, 1)  # N H W C
        bad = (e > 1e-3 * float(a.abs().max()))
        print('  bad fraction', float(bad.float().mean()))
        print('  bad per image row (h):', bad.float().mean(dim=(0, 2, 3)).cpu().numpy().round(2))
        print('  bad per w:', bad.float().mean(dim=(0, 1, 3)).cpu().numpy().round(2))
        print('  bad per 32-ch block:', bad.float().reshape(-1, co_ // 32, 32).mean(dim=(0, 2)).cpu().numpy().round(2))
    ws = co.conv_workspace(x.device)
    print('  flags nonzero:', int((ws[:256].view(torch.int32) != 0).sum()))


run(1, 28, 48, 256, 256, 3, True, True, True)
run(1, 28, 48, 256, 256, 3, True, False, True)
run(1, 28, 48, 256, 256, 3, False, False, False)
run(1, 28, 48, 256, 256, 1, False, False, False)
run(3, 7, 12, 2048, 256, 1, False, False, False)
run(1, 16, 48, 256, 256, 3, False, False, False)
run(1, 8, 32, 256, 256, 3, False, False, False)
run(4, 14, 24, 256, 256, 3, True, False, True)
run(16, 14, 24, 1024, 1024, 1, False, False, False)
