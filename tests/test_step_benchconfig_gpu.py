"""Whole-step parity AT THE BENCHED CONFIGURATION: `Model._train_on_batch` at 384x224 with exactly the code path and precision
bench.py times (tcgen05 TF32 depth CNN, bf16x3 MLP, fused re-projection, flat Adam) against oracle/step.py in fp32 on the CPU
(the restatement pinned to the reference by tests/test_oracle_step.py). Two shapes: 8 pairs at gap 1 (the bench's batch) and 2
pairs at gap 8 (the longest Euler chain; 8 pairs x 10 MLP evaluations would need ~75 GB of CPU autograd state).
Bar (BASELINE.json north_star): 1e-3 on depth maps and on the five losses. Gradients are reported by max-norm AND by share of
elements within tolerance AND by relative L2 (a 1 % bug cannot hide behind kink flips): TF32 convolutions bound what the
depth-net gradients can agree to - the reference's own GPU path (cuDNN TF32) sits at the same distance from the fp32 CPU path
(profiles/r2_probe_parity.json)."""
import pytest
import torch

from conftest import rel_err
from test_oracle_step import frac_within

pytestmark = pytest.mark.gpu
H, W = 224, 384
WATCH = ['pretrained.layer1.0.weight', 'pretrained.layer1.4.0.conv1.weight', 'pretrained.layer2.1.conv2.weight',
         'pretrained.layer3.5.conv3.weight', 'pretrained.layer4.2.bn3.weight', 'pretrained.layer4.0.downsample.0.weight',
         'scratch.layer3_rn.weight', 'scratch.refinenet2.resConfUnit1.conv1.weight', 'scratch.refinenet1.resConfUnit2.conv2.bias',
         'scratch.output_conv.0.weight', 'scratch.output_conv.2.weight', 'scratch.output_conv.4.weight']


def rel_l2(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return float(((a - b) ** 2).sum().sqrt() / (b ** 2).sum().sqrt())


@pytest.mark.timeout(900)
@pytest.mark.parametrize('n_pairs,gap', [(8, 1), (2, 8)])
def test_train_step_at_bench_configuration_matches_cpu_oracle(n_pairs, gap):
    from dvd_b200 import synthetic
    from dvd_b200.models import get_model
    from dvd_b200.networks.sceneflow_field import SceneFlowFieldNet
    from dvd_b200.third_party.MiDaS import MidasNet
    from oracle import step as ostep
    pairs = [(5 + 7 * j, 5 + 7 * j + gap) for j in range(n_pairs)]
    batch = synthetic.make_batch(pairs, H=H, W=W, seed=11 + gap)          # i.i.d. flow, as bench.py
    cpu_batch = {k: (v.squeeze(0) if torch.is_tensor(v) and v.dim() > 1 else v) for k, v in batch.items()}
    opt = synthetic.default_opt()
    depth = synthetic.seed_net_(MidasNet(non_negative=True, normalize_input=True), 0, 2000.0)
    mlp = synthetic.seed_net_(SceneFlowFieldNet(net_width=256, n_layers=4, time_dependent=True, N_freq_xyz=16, N_freq_t=16), 1)
    torch.set_num_threads(16)
    log, _, new_m, ex = ostep.train_step(depth.state_dict(), mlp.state_dict(), cpu_batch, opt, opt.warm_sf + 1)

    model = get_model('scene_flow_motion_field')(synthetic.default_opt(), None)
    model.net_depth.load_state_dict(depth.state_dict())
    model.net_sceneflow.load_state_dict(mlp.state_dict())
    model.to(torch.device('cuda:0'))
    model.net_depth.eval()
    with torch.no_grad():
        d = model.net_depth(torch.cat([cpu_batch['img_1'], cpu_batch['img_2']]).cuda())
    e1, e2 = rel_err(d[:n_pairs], ex['depth_1']), rel_err(d[n_pairs:], ex['depth_2'])
    assert e1 < 1e-3 and e2 < 1e-3, (e1, e2)
    lg = model._train_on_batch(opt.warm_sf + 1, 0, batch)
    for k in ('loss', 'flow_loss_1_2', 'disp_loss_1_2', 'sf_loss', 'acc_reg'):
        assert abs(lg[k] - log[k]) <= 1e-3 * abs(log[k]) + 1e-9, (k, lg[k], log[k])
    report = {}
    gm = dict(model.net_sceneflow.named_parameters())
    for k, ref in ex['grads_mlp'].items():
        g = gm[k].grad.reshape(ref.shape)
        report[k] = (rel_err(g, ref), frac_within(g, ref, 5e-3), rel_l2(g, ref))
        assert report[k][0] < 5e-3 and report[k][1] > 0.995 and report[k][2] < 2e-3, (k, report[k])
    gd = dict(model.net_depth.named_parameters())
    for k in WATCH:
        ref = ex['grads_depth'][k]
        g = gd[k].grad.reshape(ref.shape)
        report[k] = (rel_err(g, ref), frac_within(g, ref, 1e-2), rel_l2(g, ref))
    print('gradient report (max-norm err, share within tol, rel L2):')
    for k, v in report.items():
        print('  %-50s %.2e %.4f %.2e' % ((k,) + v))
    for k in WATCH:
        assert report[k][0] < 6e-2 and report[k][1] > 0.95 and report[k][2] < 2e-2, (k, report[k])
