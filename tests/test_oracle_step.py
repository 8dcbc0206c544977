"""CPU: the oracle's whole-step restatement (oracle/step.py, oracle/depth_nets.py) against the fixture the
reference's own `Model._train_on_batch` produced (tests/golden/step_golden.pt)."""
import os

import pytest
import torch

from conftest import GOLDEN, rel_err


@pytest.fixture(scope='module')
def step_golden():
    return torch.load(os.path.join(GOLDEN, 'step_golden.pt'), weights_only=False)


def frac_within(a, b, tol):
    """Share of elements with |a-b| <= tol * max|b| — robust to the isolated kink flips (LeakyReLU / |x| at 0)
    that make max-norm comparisons of fp32 gradients ill-conditioned."""
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() <= tol * b.abs().max()).double().mean())


def build_state(meta):
    from dvd_b200 import synthetic
    from dvd_b200.networks.sceneflow_field import SceneFlowFieldNet
    from dvd_b200.third_party.MiDaS import MidasNet
    depth = synthetic.seed_net_(MidasNet(non_negative=True, normalize_input=True), meta['depth_seed'], meta['head_bias'])
    mlp = synthetic.seed_net_(SceneFlowFieldNet(net_width=256, n_layers=4, time_dependent=True, N_freq_xyz=16, N_freq_t=16),
                              meta['mlp_seed'])
    return depth, mlp


@pytest.mark.parametrize('phase', ['warm', 'joint'])
def test_oracle_step_matches_reference(step_golden, phase):
    from dvd_b200 import synthetic
    from oracle import step
    g, meta = step_golden[phase], step_golden['meta']
    depth, mlp = build_state(meta)
    batch = synthetic.make_batch(meta['pairs'], H=meta['H'], W=meta['W'], seed=meta['batch_seed'], smooth_flow=True,
                                 flow_sigma=2.0, leading_dim=False)
    opt = synthetic.default_opt(lr=meta['lr'])
    log, new_d, new_m, ex = step.train_step(depth.state_dict(), mlp.state_dict(), batch, opt, g['epoch'])
    for k in ('loss', 'flow_loss_1_2', 'disp_loss_1_2', 'sf_loss', 'acc_reg'):
        assert abs(log[k] - g['batch_log'][k]) <= 1e-4 * abs(g['batch_log'][k]) + 1e-9, k
    assert rel_err(ex['depth_1'], g['depth_1']) < 1e-5
    assert rel_err(ex['depth_2'], g['depth_2']) < 1e-5
    assert rel_err(ex['sf_1_2'], g['sf_1_2']) < 1e-4
    for k, ref in g['mlp_grads'].items():
        assert frac_within(ex['grads_mlp'][k], ref.float(), 2e-3) > 0.999, k
    if phase == 'joint':
        for k, ref in g['depth_grads_watch'].items():
            assert frac_within(ex['grads_depth'][k], ref, 2e-3) > 0.995, k
        for k, ref in g['mlp_new'].items():
            # Adam's first step is lr * sign(g): elements whose gradient sign is ambiguous may differ by 2 lr
            assert frac_within(new_m[k], ref, 1e-6 + 2e-3 * meta['lr'] * 1000 / max(float(ref.abs().max()), 1e-9)) > 0.99, k
