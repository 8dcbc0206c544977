"""The drop-in SURFACE on a GPU: `python -m dvd_b200.train` with the reference's flags (train.py:30-364, options_train.py) ->
NetInterface.train_epoch -> DataLoader -> Model._train_on_batch, checkpoints in the reference layout, `--resume -1`, and the
GPU-resident loader (`--resident`, datasets/resident.py). Plus: a training step with the default visualisation flags reads back
exactly one small buffer per step (the reference: 5 .item() + 13 .cpu() per step)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE = ('--net scene_flow_motion_field --dataset synthetic_sequence --gaps 1,2 --n_frames 8 --height 64 --width 96 --epoch_batches 4 '
        '--lr 1e-6 --batch_size 1 --optim adam --gpu 0 --workers 0 --save_net 1 --save_net_opt --one_way --loss_type l1 --l1_mul 0 '
        '--acc_mul 1 --disp_mul 1 --warm_sf 1 --scene_lr_mul 1000 --repeat 1 --flow_mul 1 --sf_mag_div 100 --time_dependent --midas '
        '--use_disp --vis_batches_train 0 --manual_seed 1').split()


def _run(args, logdir):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    r = subprocess.run([sys.executable, '-m', 'dvd_b200.train'] + BASE + ['--logdir', logdir] + args, cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    return r.stdout


def _epochs(out):
    res = {}
    for line in out.splitlines():
        if line.startswith('epoch '):
            e = int(line.split()[1].rstrip(':'))
            res[e] = eval(line.split(':', 1)[1])      # the driver prints a plain dict of floats
    return res


@pytest.mark.timeout(1500)
def test_train_cli_checkpoints_resume_and_resident_loader(tmp_path):
    logdir = str(tmp_path / 'ckpt')
    out = _run(['--epoch', '2'], logdir)
    ep = _epochs(out)
    assert sorted(ep) == [1, 2] and all(v['loss'] == v['loss'] and v['loss'] > 0 for v in ep.values()), out[-1500:]
    full = os.path.join(logdir, 'scene_flow_motion_field_synthetic_sequence', '0')
    for f in ('nets/0001.pt', 'nets/0002.pt', 'checkpoint.pt'):
        assert os.path.exists(os.path.join(full, f)), f
    sd = torch.load(os.path.join(full, 'checkpoint.pt'), map_location='cpu', weights_only=False)
    assert set(sd) == {'nets', 'optimizers', 'epoch'} and sd['epoch'] == 2
    assert 'pretrained.layer1.0.weight' in sd['nets'][0] and 'convs.5.conv.bias' in sd['nets'][1]
    # epoch 1 is the warm-up phase (warm_sf 1): the depth net's Adam has not stepped yet; epoch 2 is joint
    assert int(float(sd['optimizers'][1]['state'][0]['step'])) == 8 and int(float(sd['optimizers'][0]['state'][0]['step'])) == 4
    # resume from the last checkpoint: one more epoch, numbered 3, Adam state carried over
    out = _run(['--epoch', '1', '--resume', '-1'], logdir)
    assert sorted(_epochs(out)) == [3], out[-1500:]
    sd = torch.load(os.path.join(full, 'checkpoint.pt'), map_location='cpu', weights_only=False)
    assert sd['epoch'] == 3 and int(float(sd['optimizers'][1]['state'][0]['step'])) == 12
    # GPU-resident sequence + gap-bucketed batches of 2 pairs
    out = _run(['--epoch', '1', '--resident', '--pairs_per_step', '2', '--epoch_batches', '3'], str(tmp_path / 'res'))
    ep = _epochs(out)
    assert sorted(ep) == [1] and ep[1]['loss'] > 0


def test_one_device_to_host_copy_per_training_step(tmp_path):
    """default visualisation flags (vis_every_train 1, vis_batches_train 10, no --vis_at_start, epoch_batches unset) + a logdir:
    nothing is dumped, and each step copies ONE 36-byte buffer to the host."""
    from torch.profiler import ProfilerActivity, profile
    from dvd_b200 import synthetic
    from dvd_b200.models import get_model
    opt = synthetic.default_opt(full_logdir=str(tmp_path), vis_every_train=1, vis_batches_train=10, vis_at_start=False, epoch_batches=None)
    model = get_model('scene_flow_motion_field')(opt, None)
    synthetic.seed_net_(model.net_depth, 0, 2000.0)
    synthetic.seed_net_(model.net_sceneflow, 1)
    model.to(torch.device('cuda:0'))
    batch = synthetic.make_batch([(3, 5)], H=64, W=96, seed=2)
    for i in range(4):      # eager steps, capture, replay
        model._train_on_batch(6, i, batch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(3):
            model._train_on_batch(6, 4 + i, batch)
        torch.cuda.synchronize()
    d2h = [e for e in prof.events() if 'memcpy' in e.name.lower() and 'dtoh' in e.name.lower()]
    assert len(d2h) == 3, [e.name for e in d2h]
    assert not os.listdir(str(tmp_path))
