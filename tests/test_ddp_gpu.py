"""GPU x2 (skipped on a single-GPU box; run with `gpurun --gpus 2`): one process per GPU over NCCL.
  * identical shards on both ranks  ==> the all-reduced / averaged step equals the single-GPU step;
  * different shards                ==> both ranks hold identical parameters after the step (flat broadcast, flat
    gradient all-reduce, same Adam update) and the gradient is the mean of the per-rank gradients."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _digest(model):
    """per-parameter (sum, abs-sum) in float64 — small enough to ship through a queue"""
    out = {}
    for i, net in enumerate(model._nets):
        for k, p in net.named_parameters():
            d = p.detach().double()
            out['%d.%s' % (i, k)] = (float(d.sum()), float(d.abs().sum()), d.numel())
    return out


def _worker(rank, world, port, same_shard, midas, q, steps=1):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from dvd_b200 import synthetic
    from dvd_b200.models import get_model
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    torch.backends.cudnn.allow_tf32 = False
    opt = synthetic.default_opt(midas=midas, lr=1e-4, multiprocess_distributed=True, global_rank=rank)
    model = get_model('scene_flow_motion_field')(opt, None)
    synthetic.seed_net_(model.net_depth, 10 + rank, 2000.0 if midas else None)       # ranks start different: the broadcast must fix that
    synthetic.seed_net_(model.net_sceneflow, 20 + rank)
    model.to(torch.device('cuda', rank))
    model.sync_parameters(0)
    pairs = [(10, 12)] if same_shard else [(10 + 7 * rank, 12 + 7 * rank)]
    batch = synthetic.make_batch(pairs, H=64, W=96, seed=3 if same_shard else 3 + rank, smooth_flow=True)
    for i in range(steps):       # steps > 3: two eager steps, the capture of the step graph (with its all-reduces), replays
        log = model._train_on_batch(6, i, batch)
    stats = dict(getattr(model, 'graph_stats', {}))
    stats['error'] = getattr(model, 'graph_error', None)
    q.put((rank, log['loss'], _digest(model), stats))
    model.release_graphs()
    dist.destroy_process_group()


def _single(midas, q, steps=1):
    sys.path.insert(0, ROOT)
    from dvd_b200 import synthetic
    from dvd_b200.models import get_model
    torch.cuda.set_device(0)
    torch.backends.cudnn.allow_tf32 = False
    opt = synthetic.default_opt(midas=midas, lr=1e-4)
    model = get_model('scene_flow_motion_field')(opt, None)
    synthetic.seed_net_(model.net_depth, 10, 2000.0 if midas else None)
    synthetic.seed_net_(model.net_sceneflow, 20)
    model.to(torch.device('cuda', 0))
    batch = synthetic.make_batch([(10, 12)], H=64, W=96, seed=3, smooth_flow=True)
    for i in range(steps):
        log = model._train_on_batch(6, i, batch)
    q.put((-1, log['loss'], _digest(model)))


def _run(target, n, args):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=target, args=(r, *args[:-1], q, args[-1]) if n > 1 else (*args[:-1], q, args[-1])) for r in range(n)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return sorted(out)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('midas', [False, True])
def test_two_rank_step_equals_single_gpu_step_on_identical_shards(midas):
    """midas=True: the tcgen05 depth engine with channels-last flat buffers and the bucketed, backward-overlapped all-reduce."""
    port = 29700 + os.getpid() % 200 + (7 if midas else 0)
    two = _run(_worker, 2, (2, port, True, midas, 1))
    one = _run(_single, 1, (midas, 1))[0]
    for r in two:
        assert abs(r[1] - one[1]) <= 1e-5 * abs(one[1])
        bad = []
        for k, (s1, a1, n) in one[2].items():
            s2 = r[2][k][0]
            # Adam's first step moves every element by ~lr = 1e-4; elements whose gradient sign is ambiguous at fp32
            # rounding level may land 2e-4 apart: tolerate 0.5 % of them
            # (MiDaS: fp32 atomics in the weight-gradient reductions make even two runs of one GPU differ in the last bits)
            if abs(s1 - s2) > (0.02 if midas else 0.005) * n * 2e-4 + 1e-6 * a1:
                bad.append((k, n, s1, s2))
        assert not bad, bad[:8]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('midas', [False, True])
def test_ranks_stay_in_lockstep_on_different_shards(midas):
    port = 29900 + os.getpid() % 90 + (5 if midas else 0)
    two = _run(_worker, 2, (2, port, False, midas, 1))
    assert two[0][1] != two[1][1]                    # different data, different losses
    assert two[0][2] == two[1][2]                    # bit-identical parameters on both ranks


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_captured_step_graph_with_all_reduce_keeps_ranks_in_lockstep():
    """5 steps of one signature on 2 ranks with different shards: eager, eager, capture (the graph contains the three bucketed NCCL
    all-reduces on NCCL's stream), replay, replay. Both ranks must have replayed and must hold bit-identical parameters."""
    port = 29500 + os.getpid() % 150
    two = _run(_worker, 2, (2, port, False, True, 5))
    for r in two:
        assert r[3].get('error') is None, r[3]
        assert r[3].get('captured') == 1 and r[3].get('replayed') == 3, r[3]
    assert two[0][1] != two[1][1]
    assert two[0][2] == two[1][2]
