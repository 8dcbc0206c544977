"""GPU: evaluation / test forward (`_predict_on_batch(is_train=False)`, `_vali_on_batch`, `test_on_batch`;
reference smf.py:265-276, video_base.py:66-103,128-155) against the CPU oracle."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def test_eval_forward_matches_oracle():
    from dvd_b200 import synthetic
    from dvd_b200.models import get_model
    from oracle import depth_nets, geometry, sf_mlp
    torch.backends.cudnn.allow_tf32 = False
    opt = synthetic.default_opt()
    model = get_model('scene_flow_motion_field')(opt, None)
    synthetic.seed_net_(model.net_depth, 0, 2000.0)
    synthetic.seed_net_(model.net_sceneflow, 1)
    sd_d = {k: v.clone() for k, v in model.net_depth.state_dict().items()}
    sd_m = {k: v.clone() for k, v in model.net_sceneflow.state_dict().items()}
    model.to(torch.device('cuda:0'))
    H, W = 64, 96
    pair = synthetic.make_batch([(12, 13)], H=H, W=W, seed=2, leading_dim=False)
    # per-frame batch of the reference's validation loader (datasets/davis_sequence.py:114-154)
    batch = {'img': pair['img_1'], 'R_1': pair['R_1'], 't_1': pair['t_1'], 'K_inv': pair['K_inv'],
             'time_stamp_1': pair['time_stamp_1'], 'time_step': pair['time_step'],
             'depth_mvs': torch.full((1, 1, H, W), 8.0), 'pair_path': ['x']}
    model.eval()
    out = model.test_on_batch(0, batch)
    with torch.no_grad():
        d = depth_nets.midas_forward(sd_d, batch['img'])
        R1, _, t1, _, _, Kinv = geometry._poses(pair)
        P = geometry.unproject(d, R1, t1, Kinv)
        sf = sf_mlp.sf_multi_step(P, batch['time_stamp_1'], float(pair['time_step']), 1, sf_mlp.layers_from_state_dict(sd_m))
    assert rel_err(torch.from_numpy(out['depth']), d) < 1e-3
    # the scene flow is the MLP evaluated on the un-projected points of the TF32 depth map: its periodic embedding (frequencies
    # up to 17) amplifies the 1.5e-4 depth difference; on the oracle's own depth map the kernel chain agrees to 1e-3
    assert rel_err(torch.from_numpy(out['sf_1_2']), sf) < 1e-2
    from dvd_b200 import ops
    with torch.no_grad():
        Pg = ops.unproject_fwd(d.cuda().contiguous(), ops.pack_poses(pair['K'], pair['K_inv'], pair['R_1_T'], pair['R_1_T'],
                                                                    pair['t_1'], pair['t_1']).cuda(), 1)
        sfg = ops.mlp_chain_fwd(model.net_sceneflow.packed(opt.sf_mag_div), Pg, batch['time_stamp_1'].cuda().contiguous(),
                                float(pair['time_step']), 1, 1, want_steps=False)['acc']
    assert rel_err(sfg, sf) < 1e-3
    log = model._vali_on_batch(1, 0, batch)
    ref = torch.nn.functional.mse_loss(1 / d, torch.full_like(d, 1 / 8.0)).item()   # (a target near the prediction would make the MSE ill-conditioned)
    assert abs(log['loss'] - ref) <= 1e-3 * abs(ref) and log['size'] == 1
