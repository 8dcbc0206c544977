"""CPU, authoring container only (skipped where /root/reference is absent, e.g. on the GPU box): the oracle
restatement and the host-side mirrors against the UNMODIFIED reference code imported through
oracle/ref_harness.py. The same comparisons, frozen, are the committed fixtures in tests/golden/."""
import pytest
import torch

from conftest import rel_err
from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason='reference tree not present')


@pytest.fixture(scope='module')
def ns():
    return ref_harness.import_reference()


def _cf(x):
    return x.squeeze(3).permute(0, 3, 1, 2)


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_geometry_matches_reference_modules(ns, dtype):
    from dvd_b200 import synthetic
    from oracle import geometry
    B, H, W = 2, 40, 56
    b = synthetic.make_batch([(3, 5), (10, 18)], H=H, W=W, dtype=dtype, leading_dim=False, flow_sigma=6.0)
    d1 = synthetic.make_depths(B, H, W, seed=1, dtype=dtype)
    d2 = synthetic.make_depths(B, H, W, seed=2, dtype=dtype)
    d1[0, 0, :3, :5] = 150.0
    d1[1, 0, 10:12, :] = -1.0
    sf = torch.randn(B, 3, H, W, dtype=dtype) * 0.05
    fb, sl = ns.sfp.flow_by_depth(), ns.sfp.scene_flow_projection_slack()
    if dtype == torch.float64:
        yy, xx = torch.meshgrid(torch.arange(H).double(), torch.arange(W).double(), indexing='ij')
        coord = torch.ones([1, H, W, 1, 3], dtype=dtype)
        coord[0, ..., 0, 0], coord[0, ..., 0, 1] = xx, yy
        fb.coord, sl.coord = coord, coord.clone()
    pose = {k: b[k] for k in ('R_1', 'R_2', 'R_1_T', 'R_2_T', 't_1', 't_2', 'K', 'K_inv')}
    r1 = fb(depth_1=d1, depth_2=d2, flow_1_2=b['flow_1_2'], **pose)
    sfl = sf.permute(0, 2, 3, 1)[..., None, :]
    r2 = sl(depth_1=d1, depth_2=d2, flow_1_2=b['flow_1_2'], flow_2_1=b['flow_2_1'], sflow_1_2=sfl, sflow_2_1=sfl, **pose)
    o = geometry.reproject(d1, d2, sf, b)
    tol = 5e-6 if dtype == torch.float32 else 1e-12
    ref = {'global_p1': _cf(r1['global_p1']), 'sf_by_depth': _cf(r1['sf_by_depth']),
           'warped_p2_camera_2': _cf(r2['warped_p2_camera_2']), 'p1_camera_2': _cf(r2['p1_camera_2']),
           'dflow_1_2': r2['dflow_1_2'].permute(0, 3, 1, 2), 'staticflow_1_2': r2['staticflow_1_2'].permute(0, 3, 1, 2),
           'depth_image_1_2': r2['depth_image_1_2'], 'depth_warp_1_2': r2['depth_warp_1_2']}
    for k, v in ref.items():
        assert rel_err(o[k], v) < tol, k


def test_depth_net_mirrors_and_functional_oracle_match_reference(ns):
    from dvd_b200 import synthetic
    from dvd_b200.third_party import MiDaS as M, hourglass as HG
    from oracle import depth_nets
    x = torch.rand(2, 3, 64, 96)
    ref = synthetic.seed_net_(ns.midas.MidasNet(path=None, non_negative=True, normalize_input=True), 0, 2000.0).eval()
    mine = synthetic.seed_net_(M.MidasNet(non_negative=True, normalize_input=True), 0, 2000.0).eval()
    assert list(ref.state_dict()) == list(mine.state_dict())
    with torch.no_grad():
        a = ref(x.clone())
        # the MiDaS mirror only holds parameters (all of its arithmetic is CUDA: depth_engine.py); on the CPU its state dict
        # drives the functional oracle, and it refuses to run itself
        assert rel_err(depth_nets.midas_forward(mine.state_dict(), x), a) < 1e-5
        assert rel_err(depth_nets.midas_forward(ref.state_dict(), x), a) < 1e-5
        import pytest
        with pytest.raises(RuntimeError):
            mine(x)
    rh = synthetic.seed_net_(ns.hourglass.HourglassModel_Embed(noexp=False), 0)
    mh = synthetic.seed_net_(HG.HourglassModel_Embed(noexp=False), 0)
    assert list(rh.state_dict()) == list(mh.state_dict())
    rh.defrost(), mh.defrost()
    with torch.no_grad():
        a = rh(x.clone())
        assert rel_err(mh(x), a) < 1e-6
        assert rel_err(depth_nets.hourglass_forward(rh.state_dict(), x), a) < 1e-5


def test_mlp_oracle_matches_reference_module(ns):
    from oracle import sf_mlp
    torch.manual_seed(3)
    net = ns.sff.SceneFlowFieldNet(net_width=256, n_layers=4, time_dependent=True, N_freq_xyz=16, N_freq_t=16)
    p = torch.randn(1, 3, 9, 11) * 3
    t = torch.full((1, 1, 9, 11), 0.4)
    with torch.no_grad():
        assert rel_err(sf_mlp.mlp_forward(p, t, sf_mlp.layers_from_state_dict(net.state_dict())), net(p, t)) < 1e-5
