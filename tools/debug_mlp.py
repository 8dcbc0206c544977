"""Decode the saved activation / dY blocks of the MLP kernels and compare each layer with a fp64
torch autograd restatement on the GPU box. Diagnostic only."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def decode_blocks(buf, off, rows, nq, npx):
    """[plane hi | plane lo] of nq interleaved MN-major blocks ([rows/8][64 px][8 ch], sf_mlp_layout.cuh:
    kActInterleave = true) -> fp32 [rows, npx]"""
    blk = (rows + 7) // 8 * 1024
    plane_bytes = nq * blk
    out = None
    for pl in range(2):
        raw = buf[off + pl * plane_bytes: off + (pl + 1) * plane_bytes].view(torch.int16).reshape(nq, rows // 8, 64, 8)
        vals = raw.view(torch.bfloat16).float().permute(1, 3, 0, 2).reshape(rows, nq * 64)[:, :npx]
        out = vals if out is None else out + vals
    return out


def main():
    from dvd_b200 import ops, _lib
    from oracle import sf_mlp
    g = torch.load(os.path.join(ROOT, 'tests', 'golden', 'mlp_golden.pt'), weights_only=False)
    sd = g['state_dict']
    ws = [sd['convs.%d.conv.weight' % l].reshape(sd['convs.%d.conv.weight' % l].shape[0], -1).cuda().contiguous() for l in range(6)]
    bs = [sd['convs.%d.conv.bias' % l].cuda().contiguous() for l in range(6)]
    cfg = ops.make_mlp_cfg()
    pk = ops.PackedMlp(cfg, 'cuda').refresh(ws, bs)
    P1, ts, dt = g['P1'].cuda(), g['ts'].cuda(), g['dt']
    B, _, H, W = P1.shape
    npx = B * H * W
    f = ops.mlp_chain_fwd(pk, P1, ts, dt, 1, 1, save=True)
    lib = _lib.load()
    # layout mirror (sf_mlp_layout.cuh)
    nin, kpad0 = 132, 144
    ntiles = (npx + 127) // 128
    nq = ntiles * 2
    rows_x = [kpad0, 256, 256, 256, 256, 256]
    xs_off, o = [], 0
    for l in range(6):
        xs_off.append(o)
        o += 2 * nq * ((rows_x[l] + 7) // 8) * 1024
    mask_off = o
    rows_dy = [256] * 5 + [16]
    dy_off, o = [], 0
    for l in range(6):
        dy_off.append(o)
        o += 2 * nq * ((rows_dy[l] + 7) // 8) * 1024
    # fp64 reference with autograd, keeping intermediates
    p = P1.double().reshape(B, 3, -1).permute(1, 0, 2).reshape(3, npx).clone().requires_grad_()
    t = ts.double().reshape(B, 1, -1).permute(1, 0, 2).reshape(1, npx)
    fx = sf_mlp.freqs(16, torch.float64).cuda()
    def emb(x):
        out = [x]
        for fn in (torch.cos, torch.sin):
            for k in range(16):
                out.append(fn(fx[k] * x))
        return torch.cat(out, 0)
    X = [torch.cat([emb(t), emb(p)], 0)]
    X[0].retain_grad()
    Ys = []
    for l in range(6):
        y = ws[l].double() @ X[-1] + bs[l].double().reshape(-1, 1)
        y.retain_grad()
        Ys.append(y)
        if l < 5:
            x = torch.nn.functional.leaky_relu(y, 0.2)
            x.retain_grad()
            X.append(x)
    s = Ys[5] / 100.0
    cot = g['cot'].cuda().double().reshape(B, 3, -1).permute(1, 0, 2).reshape(3, npx)
    (s * cot).sum().backward()
    sv = f['save']
    for l in range(6):
        mine = decode_blocks(sv, xs_off[l], rows_x[l], nq, npx).double()
        ref = X[l].detach()
        if l == 0:
            ref = torch.cat([ref, torch.zeros(kpad0 - nin, npx, device='cuda', dtype=torch.float64)], 0)
        print('X_%d err %.3e' % (l, ((mine - ref).abs().max() / ref.abs().max()).item()))
    # run dgrad only (e=0) and decode dY
    per = lib.dvd_mlp_save_bytes_per_eval(ctypes.byref(cfg), npx)
    dy = torch.zeros(lib.dvd_mlp_dy_bytes(ctypes.byref(cfg), npx), dtype=torch.uint8, device='cuda')
    a_out = torch.zeros(B, 3, H, W, device='cuda')
    gb5 = torch.zeros(3, device='cuda')
    g_acc = g['cot'].cuda().contiguous()
    P = ctypes.c_void_p
    ops._lib.check(lib.dvd_mlp_dgrad(ctypes.byref(cfg), P(pk.bwd.data_ptr()), P(f['p_steps'][0].data_ptr()), P(ts.data_ptr()),
                                     float(dt), 0, 1, P(g_acc.data_ptr()), P(0), P(0), P(a_out.data_ptr()), P(sv.data_ptr()),
                                     P(dy.data_ptr()), P(gb5.data_ptr()), npx, H * W, ops._stream()), 'dgrad')
    torch.cuda.synchronize()
    for l in range(5, -1, -1):
        mine = decode_blocks(dy, dy_off[l], rows_dy[l], nq, npx).double()
        ref = Ys[l].grad
        if l == 5:
            ref = torch.cat([ref, torch.zeros(13, npx, device='cuda', dtype=torch.float64)], 0)
        e = (mine - ref).abs()
        print('dY_%d err %.3e  (max at ch %d px %d)' % (l, (e.max() / ref.abs().max()).item(), int(e.argmax() // npx), int(e.argmax() % npx)))
        if l == 4:
            bad = (e > 1e-3 * ref.abs().max())
            print('  bad elements', int(bad.sum()), 'of', bad.numel(), '; bad per channel (first 32 ch):', bad.sum(1)[:32].tolist())
            print('  bad per pixel (first 16 px):', bad.sum(0)[:16].tolist(), ' bad px count', int((bad.sum(0) > 0).sum()))
            idxs = bad.nonzero()[:12]
            for (c, px) in idxs.tolist():
                print('   ch %d px %d mine %.6e ref %.6e ratio %.4f  y4 %.4e' % (c, px, mine[c, px].item(), ref[c, px].item(), (mine[c, px] / ref[c, px]).item(), Ys[4][c, px].item()))
            dx5 = ws[5].double().t() @ Ys[5].grad  # [256, npx]
            unmasked_err = (mine - dx5).abs()
            print('  vs unmasked dX_5: frac elements equal', float((unmasked_err < 1e-4 * dx5.abs().max()).float().mean()))
            print('  vs 0.2*dX_5: frac equal', float(((mine - 0.2 * dx5).abs() < 1e-4 * dx5.abs().max()).float().mean()))
    gx0 = X[0].grad  # embedding gradient
    gp_ref = p.grad  # [3, npx]
    mine = a_out.reshape(B, 3, -1).permute(1, 0, 2).reshape(3, npx).double()
    e = (mine - gp_ref).abs()
    print('g_p err %.3e' % (e.max() / gp_ref.abs().max()).item(), 'per-dim', [(e[d].max() / gp_ref.abs().max()).item() for d in range(3)])
    print('gb5', gb5.tolist(), 'ref', Ys[5].grad.sum(1).tolist())
    # which features matter: recompute g_p from the reference embedding gradient, piecewise
    worst = int(e.max(0).values.argmax())
    print('worst px', worst, 'mine', mine[:, worst].tolist(), 'ref', gp_ref[:, worst].tolist(), 'p', p[:, worst].tolist())
    # masks
    mk = sv[mask_off: mask_off + 5 * ntiles * 128 * 32].view(torch.int32).reshape(5, ntiles * 128, 8)
    for l in range(5):
        bits = ((mk[l].unsqueeze(-1) >> torch.arange(32, device='cuda').reshape(1, 1, 32)) & 1).reshape(ntiles * 128, 256)[:npx].t()
        ref = (Ys[l].detach() > 0)
        print('mask_%d mismatches %d' % (l, int((bits.bool() != ref).sum())))


if __name__ == '__main__':
    main()
