"""GPU: flag variants of the optimisation step against the CPU oracle (oracle/step.py, itself pinned to the
reference by tests/test_oracle_step.py): hourglass depth net, --weight_steps, --use_motion_seg, sf-loss instead of
--use_disp, --warm_reg, --use_disp_ratio."""
import pytest
import torch

from conftest import TF32_GRAD_L2, TF32_GRAD_SLOPE, grad_agreement
from test_oracle_step import frac_within

pytestmark = pytest.mark.gpu

H, W = 64, 96

VARIANTS = {
    'hourglass': dict(midas=False, lr=1e-4),
    'weight_steps_motion_seg': dict(weight_steps=True, use_motion_seg=True, lr=1e-4),
    'sf_loss': dict(use_disp=False, lr=1e-4),
    'disp_ratio_warm_reg': dict(use_disp=False, use_disp_ratio=True, warm_reg=True, lr=1e-4),
}


def _build(over):
    from dvd_b200 import synthetic
    from dvd_b200.models import get_model
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    opt = synthetic.default_opt(**over)
    model = get_model('scene_flow_motion_field')(opt, None)
    synthetic.seed_net_(model.net_depth, 0, 2000.0 if opt.midas else None)
    synthetic.seed_net_(model.net_sceneflow, 1)
    sd_d = {k: v.clone() for k, v in model.net_depth.state_dict().items()}
    sd_m = {k: v.clone() for k, v in model.net_sceneflow.state_dict().items()}
    model.to(torch.device('cuda:0'))
    return model, opt, sd_d, sd_m


@pytest.mark.parametrize('name', sorted(VARIANTS))
@pytest.mark.parametrize('epoch', [1, 6])
def test_variant_matches_oracle(name, epoch):
    from dvd_b200 import synthetic
    from oracle import step
    model, opt, sd_d, sd_m = _build(VARIANTS[name])
    batch = synthetic.make_batch([(10, 12), (30, 32)], H=H, W=W, seed=11, smooth_flow=True, flow_sigma=2.0)
    batch['motion_seg_1'] = (torch.rand(batch['motion_seg_1'].shape, generator=torch.Generator().manual_seed(5)) > 0.3).float()
    ob = {k: (v.squeeze(0) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in batch.items()}
    log_o, new_d, new_m, ex = step.train_step(sd_d, sd_m, ob, opt, epoch)
    log = model._train_on_batch(epoch, 0, batch)
    for k in ('loss', 'flow_loss_1_2', 'disp_loss_1_2', 'sf_loss', 'acc_reg'):
        assert abs(log[k] - log_o[k]) <= 1e-3 * abs(log_o[k]) + 1e-9, (name, epoch, k, log[k], log_o[k])
    grads = dict(model.net_sceneflow.named_parameters())
    for k, ref in ex['grads_mlp'].items():
        assert frac_within(grads[k].grad.reshape(ref.shape), ref, 5e-3) > 0.99, (name, k)
    if epoch > opt.warm_sf:
        dg = dict(model.net_depth.named_parameters())
        checked = 0
        for k, ref in list(ex['grads_depth'].items())[::37]:
            if opt.midas:      # TF32 tensor-core convolutions
                slope, l2, mx = grad_agreement(dg[k].grad, ref)
                assert abs(slope) < TF32_GRAD_SLOPE and l2 < TF32_GRAD_L2, (name, k, slope, l2, mx)
            else:              # hourglass: fp32 library convolutions
                assert frac_within(dg[k].grad, ref, 1e-2) > 0.98, (name, k)
            checked += 1
        assert checked >= 3
