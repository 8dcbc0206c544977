"""Micro-benchmark of the tcgen05 scene-flow MLP kernels (CUDA events). FLOP accounting: fwd eval
593 408 FLOP/px; dgrad the same; wgrad the same (useful fp32-equivalent FLOPs, the
bf16x3 split issues 3x as many tensor-core MACs)."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def main(B=2, H=224, W=384, n_eval=2):
    from dvd_b200 import ops, _lib
    from oracle import sf_mlp
    layers = sf_mlp.init_layers(seed=1)
    ws = [w.cuda().contiguous() for w, _ in layers]
    bs = [b.cuda().contiguous() for _, b in layers]
    cfg = ops.make_mlp_cfg()
    pk = ops.PackedMlp(cfg, 'cuda').refresh(ws, bs)
    p = (torch.randn(B, 3, H, W) * 3).cuda()
    t = torch.full((B, 1, H, W), 0.25).cuda()
    npx = B * H * W
    F = 593408.0 * npx
    res = {}
    tm = ev_time(lambda: ops.mlp_chain_fwd(pk, p, t, 1 / 80, n_eval, n_eval, save=False, want_steps=False))
    res['fwd_infer'] = {'ms_per_eval': tm / n_eval * 1e3, 'TFLOPs_fp32_equiv': F * n_eval / tm / 1e12}
    f = ops.mlp_chain_fwd(pk, p, t, 1 / 80, n_eval, n_eval, save=True)
    lib = _lib.load()
    per = lib.dvd_mlp_save_bytes_per_eval(ctypes.byref(cfg), npx)
    sv = f['save']
    p_steps = f['p_steps']
    s_steps = torch.empty_like(p_steps)
    acc = torch.empty_like(p)
    P = ctypes.c_void_p

    def fwd_train():
        _lib.check(lib.dvd_mlp_chain_fwd(ctypes.byref(cfg), P(pk.fwd.data_ptr()), P(pk.bias.data_ptr()), P(p.data_ptr()),
                                         P(t.data_ptr()), 1 / 80, n_eval, n_eval, P(acc.data_ptr()), P(s_steps.data_ptr()),
                                         P(p_steps.data_ptr()), P(sv.data_ptr()), npx, H * W, ops._stream()), 'fwd')
    tm = ev_time(fwd_train)
    res['fwd_train'] = {'ms_per_eval': tm / n_eval * 1e3, 'TFLOPs_fp32_equiv': F * n_eval / tm / 1e12,
                        'save_GBps': per * n_eval / tm / 1e9}
    dy = torch.empty(lib.dvd_mlp_dy_bytes(ctypes.byref(cfg), npx), dtype=torch.uint8, device='cuda')
    g = torch.randn_like(p)
    a_out = torch.empty_like(p)
    gb5 = torch.zeros(3, device='cuda')

    def dgrad():
        _lib.check(lib.dvd_mlp_dgrad(ctypes.byref(cfg), P(pk.bwd.data_ptr()), P(p_steps[0].data_ptr()), P(t.data_ptr()),
                                     1 / 80, 0, 1, P(g.data_ptr()), P(0), P(0), P(a_out.data_ptr()), P(sv.data_ptr()),
                                     P(dy.data_ptr()), P(gb5.data_ptr()), npx, H * W, ops._stream()), 'dgrad')
    tm = ev_time(dgrad)
    res['dgrad'] = {'ms_per_eval': tm * 1e3, 'TFLOPs_fp32_equiv': F / tm / 1e12}
    gw = [torch.zeros_like(w) for w in ws]
    gb = [torch.zeros_like(b) for b in bs]
    gwa, gba = ops._ptr_array(gw), ops._ptr_array(gb)

    def wgrad():
        _lib.check(lib.dvd_mlp_wgrad(ctypes.byref(cfg), P(sv.data_ptr()), P(dy.data_ptr()), gwa, gba, npx, ops._stream()), 'wgrad')
    tm = ev_time(wgrad)
    res['wgrad'] = {'ms_per_eval': tm * 1e3, 'TFLOPs_fp32_equiv': F / tm / 1e12,
                    'operand_GBps': (per + dy.numel()) / tm / 1e9}
    tm = ev_time(lambda: pk.refresh(ws, bs))
    res['pack_weights'] = {'ms': tm * 1e3}
    for k, v in res.items():
        print(k, v, flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump({'B': B, 'H': H, 'W': W, 'n_eval': n_eval, 'results': res},
              open(os.path.join(ROOT, 'gpurun_out', 'bench_mlp.json'), 'w'), indent=1)


if __name__ == '__main__':
    main(B=int(sys.argv[1]) if len(sys.argv) > 1 else 2)
