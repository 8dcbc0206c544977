"""TEST INFRASTRUCTURE — not product code.

Imports the UNMODIFIED reference implementation from /root/reference (authoring
container only; the tree does not exist on the GPU box) so that

  * the CPU restatement in oracle/*.py can be pinned against the reference's own code, and
  * golden fixtures under tests/golden/ can be generated (oracle/gen_golden.py).

Five import shims are needed; none touches arithmetic:
  1. matplotlib stub modules (models/video_base.py:20-25 imports it at module top);
  2. inspect.getargspec alias (removed in py3.11; models/scene_flow_motion_field.py:128,134);
  3. torch.hub ResNeXt101-WSL → torchvision resnext101_32x8d(weights=None), same architecture
     (third_party/midas_blocks.py:48-50);
  4. no checkpoint files → midas_pretrain_path = None (configs/__init__.py:15-16);
  5. HTMLVisualizer → no-op (models/scene_flow_motion_field.py:124).
"""
import argparse
import inspect
import os
import sys
import types

REF_ROOT = os.environ.get('DVD_REFERENCE', '/root/reference')


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, 'losses'))


def _install_shims():
    if 'matplotlib' not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType('matplotlib')
            plt = types.ModuleType('matplotlib.pyplot')
            cm = types.ModuleType('matplotlib.cm')
            colors = types.ModuleType('matplotlib.colors')

            class ListedColormap:  # noqa: D401 - stub
                def __init__(self, *a, **k):
                    pass
            colors.ListedColormap = ListedColormap
            mpl.pyplot, mpl.cm, mpl.colors = plt, cm, colors
            sys.modules.update({'matplotlib': mpl, 'matplotlib.pyplot': plt,
                                'matplotlib.cm': cm, 'matplotlib.colors': colors})
    if not hasattr(inspect, 'getargspec'):
        inspect.getargspec = inspect.getfullargspec


_REF_MODULE_ROOTS = ('models', 'losses', 'networks', 'third_party', 'configs', 'loggers',
                     'visualize', 'util', 'datasets', 'options')


def import_reference():
    """Return a namespace with the reference modules on the hot path."""
    if not reference_available():
        raise RuntimeError('reference tree not found at %s' % REF_ROOT)
    _install_shims()
    # our own repo has same-named mirror packages inside dynamic-video-depth_b200/; make sure the
    # reference's top-level names resolve to /root/reference.
    for name in list(sys.modules):
        if name.split('.')[0] in _REF_MODULE_ROOTS:
            mod = sys.modules[name]
            f = getattr(mod, '__file__', '') or ''
            if not f.startswith(REF_ROOT):
                del sys.modules[name]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import torchvision
    import third_party.midas_blocks as mb
    mb._make_pretrained_resnext101_wsl = lambda use_pretrained: mb._make_resnet_backbone(
        torchvision.models.resnext101_32x8d(weights=None))
    import losses.scene_flow_projection as sfp
    import networks.sceneflow_field as sff
    import networks.blocks as blocks
    import third_party.MiDaS as midas
    import third_party.hourglass as hourglass
    import models.scene_flow_motion_field as smf
    smf.midas_pretrain_path = None

    class _NoVis:
        def __init__(self, *a, **k):
            pass

        def visualize(self, *a, **k):
            pass
    smf.Visualizer = _NoVis
    ns = types.SimpleNamespace(sfp=sfp, sff=sff, blocks=blocks, midas=midas, hourglass=hourglass,
                               smf=smf, mb=mb)
    return ns


class _NullLogger:
    def add_logger(self, l):
        pass

    def get_html_logger(self):
        return None


def default_opt(**over):
    """Namespace with the flags of experiments/davis/train_sequence.sh:24-63."""
    d = dict(optim='adam', lr=1e-6, adam_beta1=0.5, adam_beta2=0.9, full_logdir='/tmp/dvd_ref_log',
             global_rank=0, dataset='davis_sequence', batch_size=1, epoch_batches=2000,
             vis_every_train=10**9, vis_at_start=True, vis_batches_train=0,
             multiprocess_distributed=False,
             l1_mul=0.0, disp_mul=1.0, one_way=True, loss_type='l1', scene_lr_mul=1000.0, n_down=3,
             weight_steps=False, sf_min_mul=0, sf_quantile=0.5, static=False, static_mul=1,
             flow_mul=1.0, acc_mul=1.0, si_mul=0, cos_mul=0, motion_seg_hard=False, warm_mul=1,
             interp_steps=5, warm_static=False, use_disp=True, use_disp_ratio=False,
             time_dependent=True, use_cnn=False, use_embedding=False, use_motion_seg=False,
             warm_reg=False, warm_sf=5, n_freq_xyz=16, n_freq_t=16, sf_mag_div=100.0, midas=True)
    d.update(over)
    return argparse.Namespace(**d)


def build_reference_model(opt=None, seed=0, head_bias=2000.0, device='cpu'):
    """Reference `Model` with seeded random weights (MiDaS head bias shifted so
    that depth≈5; otherwise relu(out)=0 → depth=1e6 → mask ≡ 0 → loss ≡ 0)."""
    import torch
    ns = import_reference()
    opt = opt or default_opt()
    torch.manual_seed(seed)
    if not opt.midas:
        # hourglass: the ctor torch.load()s a checkpoint that does not exist offline
        # (models/scene_flow_motion_field.py:121); bypass the load only.
        real_load = torch.load
        torch.load = lambda *a, **k: ns.smf.HourglassModel_Embed(noexp=False).net_depth.state_dict()
        try:
            model = ns.smf.Model(opt, _NullLogger())
        finally:
            torch.load = real_load
    else:
        model = ns.smf.Model(opt, _NullLogger())
        with torch.no_grad():
            model.net_depth.scratch.output_conv[4].bias.fill_(head_bias)
    model.to(torch.device(device))
    return model, ns
