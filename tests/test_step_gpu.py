"""GPU parity of the whole optimisation step: dvd_b200 `Model._train_on_batch` (reference plug-in surface,
CUDA kernels through the C ABI, cuDNN depth net) vs the fixture produced by the reference's own
`Model._train_on_batch` on the same seeded batch / weights (tests/golden/step_golden.pt).
Tolerance: 1e-3 on losses and depth maps (BASELINE.json north_star)."""
import os

import pytest
import torch

from conftest import GOLDEN, TF32_GRAD_L2, TF32_GRAD_SLOPE, grad_agreement, rel_err
from test_oracle_step import build_state, frac_within

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def step_golden():
    return torch.load(os.path.join(GOLDEN, 'step_golden.pt'), weights_only=False)


def make_model(meta, **over):
    from dvd_b200 import synthetic
    from dvd_b200.models import get_model
    torch.backends.cudnn.allow_tf32 = False       # (only the hourglass variant still touches cuDNN; MiDaS runs the repo's TF32 kernels)
    torch.backends.cuda.matmul.allow_tf32 = False
    opt = synthetic.default_opt(lr=meta['lr'], **over)
    model = get_model('scene_flow_motion_field')(opt, None)
    depth, mlp = build_state(meta)
    model.net_depth.load_state_dict(depth.state_dict())
    model.net_sceneflow.load_state_dict(mlp.state_dict())
    model.to(torch.device('cuda:0'))
    return model


@pytest.mark.parametrize('phase', ['warm', 'joint'])
def test_train_step_matches_reference_fixture(step_golden, phase):
    from dvd_b200 import synthetic
    g, meta = step_golden[phase], step_golden['meta']
    model = make_model(meta)
    batch = synthetic.make_batch(meta['pairs'], H=meta['H'], W=meta['W'], seed=meta['batch_seed'], smooth_flow=True,
                                 flow_sigma=2.0)
    model.net_depth.eval()   # BN is in eval mode on this path (smf.py:157,168)
    with torch.no_grad():
        d1 = model.net_depth(batch['img_1'][0].cuda())
    assert rel_err(d1, g['depth_1']) < 1e-3
    log = model._train_on_batch(g['epoch'], 0, batch)
    assert set(log) >= {'size', 'loss', 'total_loss', 'flow_loss_1_2', 'disp_loss_1_2', 'sf_loss', 'acc_reg'}
    for k in ('loss', 'flow_loss_1_2', 'disp_loss_1_2', 'sf_loss', 'acc_reg'):
        assert abs(log[k] - g['batch_log'][k]) <= 1e-3 * abs(g['batch_log'][k]) + 1e-9, (k, log[k], g['batch_log'][k])
    grads = dict(model.net_sceneflow.named_parameters())
    for k, ref in g['mlp_grads'].items():
        assert frac_within(grads[k].grad.reshape(ref.shape), ref.float(), 5e-3) > 0.995, k
    if phase == 'joint':
        dg = dict(model.net_depth.named_parameters())
        for k, ref in g['depth_grads_watch'].items():
            slope, l2, mx = grad_agreement(dg[k].grad, ref)
            assert abs(slope) < TF32_GRAD_SLOPE and l2 < TF32_GRAD_L2, (k, slope, l2, mx)
        sd = model.net_sceneflow.state_dict()
        for k, ref in g['mlp_new'].items():
            tol = 1e-6 + 2e-3 * meta['lr'] * 1000 / max(float(ref.abs().max()), 1e-9)
            assert frac_within(sd[k].reshape(ref.shape), ref, tol) > 0.99, k
        sdd = model.net_depth.state_dict()
        for k, ref in g['depth_new_watch'].items():
            # Adam's first step is lr * sign(g): TF32 noise flips the sign of the smallest gradients (a 2 lr difference)
            tol = 1e-6 + 2e-3 * meta['lr'] / max(float(ref.abs().max()), 1e-9)
            assert frac_within(sdd[k], ref, tol) > 0.9, (k, frac_within(sdd[k], ref, tol))
    else:
        # warm-up: the depth net must be untouched
        depth0, _ = build_state(meta)
        for k, v in depth0.state_dict().items():
            assert torch.equal(model.net_depth.state_dict()[k].cpu(), v), k


def test_two_steps_and_checkpoint_roundtrip(step_golden, tmp_path):
    """Second step runs with re-packed weights; checkpoint layout {'nets','optimizers','epoch'} round-trips."""
    from dvd_b200 import synthetic
    meta = step_golden['meta']
    model = make_model(meta)
    batch = synthetic.make_batch([(4, 6), (20, 21)], H=64, W=96, seed=3, smooth_flow=True)
    l1 = model._train_on_batch(6, 0, batch)
    l2 = model._train_on_batch(6, 1, batch)
    assert all(map(lambda v: v == v, l2.values()))          # no NaN
    assert l2['loss'] != l1['loss']                           # parameters moved
    f = str(tmp_path / 'ckpt.pt')
    model.save_state_dict(f, save_optimizer=True, additional_values={'epoch': 6})
    sd = torch.load(f, weights_only=False)
    assert set(sd) == {'nets', 'optimizers', 'epoch'} and len(sd['nets']) == 2
    assert 'pretrained.layer1.0.weight' in sd['nets'][0] and 'convs.5.conv.bias' in sd['nets'][1]
    assert 'exp_avg' in sd['optimizers'][1]['state'][0]
    model2 = make_model(meta)
    extra = model2.load_state_dict(f)
    assert extra == {'epoch': 6}
    l3a = model._train_on_batch(6, 2, batch)
    l3b = model2._train_on_batch(6, 2, batch)
    assert abs(l3a['loss'] - l3b['loss']) <= 1e-5 * abs(l3a['loss'])


def test_operator_level_mirrors(reproject_golden):
    """losses/scene_flow_projection mirrors: same keys / shapes / values as the reference modules."""
    from dvd_b200.losses import scene_flow_projection as sfp
    g = reproject_golden
    i = g['inputs']
    b = {k: v.cuda() for k, v in i['batch'].items()}
    pose = {k: b[k] for k in ('R_1', 'R_2', 'R_1_T', 'R_2_T', 't_1', 't_2', 'K', 'K_inv')}
    d1, d2, sf = i['d1'].cuda(), i['d2'].cuda(), i['sf'].cuda()
    r1 = sfp.flow_by_depth()(depth_1=d1, depth_2=d2, flow_1_2=b['flow_1_2'], **pose)
    sfl = sf.permute(0, 2, 3, 1)[..., None, :]
    r2 = sfp.scene_flow_projection_slack()(depth_1=d1, depth_2=d2, flow_1_2=b['flow_1_2'], flow_2_1=b['flow_2_1'],
                                           sflow_1_2=sfl, sflow_2_1=sfl, **pose)
    B, _, H, W = d1.shape
    assert r1['global_p1'].shape == (B, H, W, 1, 3) and r2['dflow_1_2'].shape == (B, H, W, 2)
    assert set(r2) == {'dflow_1_2', 'depth_image_1_2', 'depth_warp_1_2', 'depth_1', 'depth_2', 'scenef_1_2',
                       'global_p1', 'staticflow_1_2', 'p1_camera_2', 'warped_p2_camera_2'}
    cf = lambda x: x.squeeze(3).permute(0, 3, 1, 2)  # noqa: E731
    assert rel_err(cf(r1['sf_by_depth']), g['tensors']['sf_by_depth']) < 5e-5
    assert rel_err(cf(r2['p1_camera_2']), g['tensors']['p1_camera_2']) < 5e-5
    assert rel_err(r2['dflow_1_2'].permute(0, 3, 1, 2), g['tensors']['dflow_1_2']) < 5e-5
    P = sfp.unproject_ptcld()(d1, b['R_1'], b['t_1'], b['K_inv'])
    assert rel_err(cf(P), g['tensors']['global_p1']) < 5e-5
