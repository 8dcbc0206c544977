"""TEST INFRASTRUCTURE — full-step golden fixture: runs the UNMODIFIED reference
`Model._train_on_batch` (models/scene_flow_motion_field.py:152-227) on a seeded synthetic batch at
64x96 with name-seeded weights (dvd_b200.synthetic.seed_net_), warm-up and joint phase, and stores the
batch_log, the depth maps, gradient / parameter-update digests. Authoring container only."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
H, W = 64, 96
PAIRS = [(10, 13)]          # gap 3 -> 3 Euler steps (+ the regulariser's two evaluations)
WATCH = ['pretrained.layer1.0.weight', 'pretrained.layer2.1.conv2.weight', 'pretrained.layer4.2.bn3.weight',
         'scratch.output_conv.2.weight', 'scratch.refinenet1.resConfUnit2.conv2.bias', 'scratch.output_conv.4.weight']


def run(midas=True):
    from dvd_b200 import synthetic
    out = {}
    for phase, epoch in (('warm', 1), ('joint', 6)):
        opt = ref_harness.default_opt(midas=midas, lr=1e-4)   # larger lr than the experiments so that the update is visible in fp32
        model, ns = ref_harness.build_reference_model(opt, seed=0)
        synthetic.seed_net_(model.net_depth, 0, 2000.0 if midas else None)
        synthetic.seed_net_(model.net_sceneflow, 1)
        batch = synthetic.make_batch(PAIRS, H=H, W=W, seed=7, smooth_flow=True, flow_sigma=2.0)
        sd_d0 = {k: v.clone() for k, v in model.net_depth.state_dict().items()}
        sd_m0 = {k: v.clone() for k, v in model.net_sceneflow.state_dict().items()}
        keep = {}
        orig = model._calc_loss

        def spy(pred, _o=orig, _k=keep):
            _k['depth_1'] = pred['depth_1'].detach().clone()
            _k['depth_2'] = pred['depth_2'].detach().clone()
            _k['sf_1_2'] = pred['sf_1_2'].detach().clone()
            return _o(pred)
        model._calc_loss = spy
        log = model._train_on_batch(epoch, 0, {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()})
        sd_d1, sd_m1 = model.net_depth.state_dict(), model.net_sceneflow.state_dict()
        grads_d = {k: (p.grad.clone() if p.grad is not None else None) for k, p in model.net_depth.named_parameters()}
        grads_m = {k: p.grad.clone() for k, p in model.net_sceneflow.named_parameters()}
        out[phase] = {
            'epoch': epoch, 'batch_log': {k: float(v) for k, v in log.items()},
            'depth_1': keep['depth_1'], 'depth_2': keep['depth_2'], 'sf_1_2': keep['sf_1_2'],
            'mlp_grads': {k: v.half() if phase == 'warm' else v for k, v in grads_m.items()},
            'mlp_new': {k: v.clone() for k, v in sd_m1.items()} if phase == 'joint' else {},
            'depth_grads_watch': {k: grads_d[k] for k in WATCH if grads_d.get(k) is not None},
            'depth_new_watch': {k: sd_d1[k].clone() for k in WATCH},
            'depth_grad_norms': {k: float(g.norm()) for k, g in grads_d.items() if g is not None},
            'depth_delta_abs_sum': {k: float((sd_d1[k] - sd_d0[k]).abs().sum()) for k in sd_d0 if sd_d0[k].dtype.is_floating_point},
        }
        print(phase, out[phase]['batch_log'])
    out['meta'] = {'H': H, 'W': W, 'pairs': PAIRS, 'batch_seed': 7, 'lr': 1e-4, 'midas': midas,
                   'depth_seed': 0, 'mlp_seed': 1, 'head_bias': 2000.0}
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    out = run(True)
    torch.save(out, os.path.join(GOLD, 'step_golden.pt'))
    print('wrote step_golden.pt (%.1f MB)' % (os.path.getsize(os.path.join(GOLD, 'step_golden.pt')) / 1e6))


if __name__ == '__main__':
    main()
