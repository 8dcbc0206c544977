"""Section timing of one optimisation step (CUDA events): where does the step time go?"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ev(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main(B=4, gap=4):
    from dvd_b200 import ops, synthetic
    from dvd_b200.models import get_model
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    H, W = 224, 384
    opt = synthetic.default_opt()
    model = get_model('scene_flow_motion_field')(opt, None)
    synthetic.seed_net_(model.net_depth, 0, 2000.0)
    synthetic.seed_net_(model.net_sceneflow, 1)
    model.to(torch.device('cuda:0'))
    pairs = [(3 * j, 3 * j + gap) for j in range(B)]
    hb = synthetic.make_batch(pairs, H=H, W=W, seed=0)
    rb = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in hb.items()}
    rb['time_step'] = hb['time_step']
    rb['steps_hint'] = gap
    res = {}
    res['full_step_ms'] = ev(lambda: model._train_on_batch(6, 0, rb))
    # sections
    b = {k: (v.squeeze(0) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in rb.items()}
    img = torch.cat([b['img_1'], b['img_2']], 0)
    model._set_depth_trainable(True)

    def depth_fb():
        d = model.net_depth(img)
        d.sum().backward()
    res['depth_fwd_bwd_ms'] = ev(depth_fb)
    with torch.no_grad():
        res['depth_fwd_ms'] = ev(lambda: model.net_depth(img))
        d = model.net_depth(img)
    d1, d2 = d[:B].contiguous(), d[B:].contiguous()
    poses = ops.pack_poses(b['K'], b['K_inv'], b['R_1_T'], b['R_2_T'], b['t_1'], b['t_2'])
    P1 = ops.unproject_fwd(d1, poses, 1)
    ts1 = b['time_stamp_1'].contiguous()
    n_eval = max(gap, 2)
    net = model.net_sceneflow
    pk = net.packed(opt.sf_mag_div)
    ws = [w.reshape(w.shape[0], -1) for w in net.weights()]

    def mlp_fb():
        p = P1.clone().requires_grad_()
        acc, st = ops.scene_flow_chain(p, ts1, pk, 1 / 80, n_eval, gap, ws, net.biases())
        acc.sum().backward()
    res['mlp_chain_fwd_bwd_ms'] = ev(mlp_fb)
    res['mlp_chain_fwd_infer_ms'] = ev(lambda: ops.mlp_chain_fwd(pk, P1, ts1, 1 / 80, n_eval, gap, want_steps=False))
    sf = ops.mlp_chain_fwd(pk, P1, ts1, 1 / 80, n_eval, gap)['acc']
    mask = b['mask_2'].reshape(B, H, W).contiguous()
    cfg = model._loss_cfg()
    flow = b['flow_1_2'].contiguous()

    def reproj():
        s = ops.reproject_loss_fwd(d1, d2, flow, mask, sf, poses, cfg)
        ops.reproject_loss_bwd(d1, d2, flow, mask, sf, poses, cfg, s)
    res['reproject_fwd_bwd_ms'] = ev(reproj)

    def adam():
        model.optimizer_depth.step()
        model.optimizer_scene.step()
    res['adam_ms'] = ev(adam)
    res['pack_ms'] = ev(lambda: net.packed(opt.sf_mag_div, force=True))
    res['config'] = {'B': B, 'gap': gap, 'n_eval': n_eval}
    for k, v in res.items():
        print(k, v, flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'profile_step.json'), 'w'), indent=1)


if __name__ == '__main__':
    main(B=int(sys.argv[1]) if len(sys.argv) > 1 else 4, gap=int(sys.argv[2]) if len(sys.argv) > 2 else 4)
