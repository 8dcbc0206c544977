#!/usr/bin/env python
"""Round-2 first probe (GPU box): (1) whole-step parity of the product path at the BENCHED configuration (384x224,
bench.py's precision flags) against oracle/step.py on the CPU, for several precision settings of the depth-net convolutions;
(2) the reference-equivalent eager PyTorch step (oracle/step.py moved to cuda:0, torch defaults) timed = the denominator of
the >=10x target. Writes gpurun_out/r2_probe_parity.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402

H, W = 224, 384


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def frac_within(a, b, tol):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() <= tol * b.abs().max()).double().mean())


def main():
    from dvd_b200 import synthetic
    from dvd_b200.models import get_model
    from dvd_b200.networks.sceneflow_field import SceneFlowFieldNet
    from dvd_b200.third_party import MiDaS
    from dvd_b200.third_party.MiDaS import MidasNet
    from oracle import step as ostep
    out = {}
    pairs = [(10, 14), (30, 34)]
    batch = synthetic.make_batch(pairs, H=H, W=W, seed=5)
    cpu_batch = {k: (v.squeeze(0) if torch.is_tensor(v) and v.dim() > 1 else v) for k, v in batch.items()}
    opt = synthetic.default_opt()
    depth = synthetic.seed_net_(MidasNet(non_negative=True, normalize_input=True), 0, 2000.0)
    mlp = synthetic.seed_net_(SceneFlowFieldNet(net_width=256, n_layers=4, time_dependent=True, N_freq_xyz=16, N_freq_t=16), 1)
    torch.set_num_threads(16)
    t0 = time.time()
    log, new_d, new_m, ex = ostep.train_step(depth.state_dict(), mlp.state_dict(), cpu_batch, opt, 6)
    out['cpu_oracle_seconds'] = time.time() - t0
    out['cpu_log'] = {k: v for k, v in log.items()}
    watch = ['pretrained.layer1.4.0.conv1.weight', 'pretrained.layer2.1.conv2.weight', 'pretrained.layer3.5.conv3.weight',
             'pretrained.layer4.2.bn3.weight', 'scratch.refinenet2.resConfUnit1.conv1.weight', 'scratch.output_conv.0.weight',
             'pretrained.layer1.0.weight']

    def run(name, tf32, tc_train):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        MiDaS._TC_CONV_TRAIN = tc_train
        model = get_model('scene_flow_motion_field')(synthetic.default_opt(), None)
        model.net_depth.load_state_dict(depth.state_dict())
        model.net_sceneflow.load_state_dict(mlp.state_dict())
        model.to(torch.device('cuda:0'))
        model.net_depth.eval()
        r = {}
        try:
            with torch.no_grad():
                d1 = model.net_depth(cpu_batch['img_1'].cuda())
            r['depth_1'] = rel_err(d1, ex['depth_1'])
            lg = model._train_on_batch(6, 0, batch)
            torch.cuda.synchronize()
            r['log'] = {k: (lg[k], log[k], abs(lg[k] - log[k]) / max(abs(log[k]), 1e-12)) for k in
                        ('loss', 'flow_loss_1_2', 'disp_loss_1_2', 'sf_loss', 'acc_reg')}
            gm = dict(model.net_sceneflow.named_parameters())
            r['mlp_grads'] = {k: (rel_err(gm[k].grad.reshape(v.shape), v), frac_within(gm[k].grad.reshape(v.shape), v, 5e-3))
                              for k, v in ex['grads_mlp'].items()}
            gd = dict(model.net_depth.named_parameters())
            r['depth_grads'] = {k: (rel_err(gd[k].grad, ex['grads_depth'][k]), frac_within(gd[k].grad, ex['grads_depth'][k], 5e-3),
                                    float((gd[k].grad.double().cpu() * ex['grads_depth'][k].double()).sum() /
                                          (ex['grads_depth'][k].double() ** 2).sum()) - 1.0) for k in watch}
        except Exception as e:   # noqa: BLE001
            r['error'] = repr(e)[:500]
        out[name] = r
        print(name, json.dumps(r), flush=True)
        del model
        torch.cuda.empty_cache()

    run('cudnn_fp32', False, False)
    run('cudnn_tf32', True, False)
    if os.environ.get('PROBE_TC', '1') == '1':
        run('tcgen05_train_tf32', True, True)
    MiDaS._TC_CONV_TRAIN = False

    # ---- (2) eager PyTorch reference-equivalent step on cuda:0 (torch defaults: cuDNN TF32 on, matmul fp32) ----
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = False
    dev = torch.device('cuda:0')
    sd_d = {k: v.to(dev) for k, v in depth.state_dict().items()}
    sd_m = {k: v.to(dev) for k, v in mlp.state_dict().items()}

    def gpu_ref(pairs_list, n_warm=2):
        bs = []
        for i, pr in enumerate(pairs_list):
            b = synthetic.make_batch(pr, H=H, W=W, seed=100 + i, leading_dim=False)
            bs.append({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()})
        ad, am = {}, {}
        d, m = sd_d, sd_m
        for i in range(n_warm):
            _, d, m, _ = ostep.train_step(d, m, bs[i % len(bs)], opt, 6, ad, am)
        torch.cuda.synchronize()
        t0 = time.time()
        n = 0
        for b in bs:
            _, d, m, _ = ostep.train_step(d, m, b, opt, 6, ad, am)
            n += b['img_1'].shape[0]
        torch.cuda.synchronize()
        return n / (time.time() - t0)

    try:
        out['gpu_eager_B1_gap2_pairs_per_s'] = gpu_ref([[(10 + i, 12 + i)] for i in range(10)])
        gaps = (8, 6, 4, 2, 1)
        out['gpu_eager_B8_gapmix_pairs_per_s'] = gpu_ref([[(3 * j + s, 3 * j + s + gaps[s % 5]) for j in range(8)] for s in range(5)])
        out['gpu_eager_B1_gapmix_pairs_per_s'] = gpu_ref([[(10 + s, 10 + s + gaps[s % 5])] for s in range(10)])
    except Exception as e:   # noqa: BLE001
        out['gpu_eager_error'] = repr(e)[:500]
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'r2_probe_parity.json'), 'w'), indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
