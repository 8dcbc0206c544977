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


def _worker(rank, world, port, same_shard, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from dvd_b200 import synthetic
    from dvd_b200.models import get_model
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    torch.backends.cudnn.allow_tf32 = False
    opt = synthetic.default_opt(midas=False, lr=1e-4, multiprocess_distributed=True, global_rank=rank)
    model = get_model('scene_flow_motion_field')(opt, None)
    synthetic.seed_net_(model.net_depth, 10 + rank)       # ranks start different: the broadcast must fix that
    synthetic.seed_net_(model.net_sceneflow, 20 + rank)
    model.to(torch.device('cuda', rank))
    model.sync_parameters(0)
    pairs = [(10, 12)] if same_shard else [(10 + 7 * rank, 12 + 7 * rank)]
    batch = synthetic.make_batch(pairs, H=64, W=96, seed=3 if same_shard else 3 + rank, smooth_flow=True)
    log = model._train_on_batch(6, 0, batch)
    flat = torch.cat([model.optimizer_depth.flat.data, model.optimizer_scene.flat.data]).double()
    q.put((rank, log['loss'], float(flat.sum()), float(flat.abs().sum()), float((flat * torch.arange(flat.numel(), device=flat.device) % 7).sum())))
    dist.destroy_process_group()


def _single(q):
    sys.path.insert(0, ROOT)
    from dvd_b200 import synthetic
    from dvd_b200.models import get_model
    torch.cuda.set_device(0)
    torch.backends.cudnn.allow_tf32 = False
    opt = synthetic.default_opt(midas=False, lr=1e-4)
    model = get_model('scene_flow_motion_field')(opt, None)
    synthetic.seed_net_(model.net_depth, 10)
    synthetic.seed_net_(model.net_sceneflow, 20)
    model.to(torch.device('cuda', 0))
    batch = synthetic.make_batch([(10, 12)], H=64, W=96, seed=3, smooth_flow=True)
    log = model._train_on_batch(6, 0, batch)
    flat = torch.cat([model.optimizer_depth.flat.data, model.optimizer_scene.flat.data]).double()
    q.put((-1, log['loss'], float(flat.sum()), float(flat.abs().sum()), float((flat * torch.arange(flat.numel(), device=flat.device) % 7).sum())))


def _run(target, n, args):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=target, args=(r, *args, q) if n > 1 else (q,)) for r in range(n)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return sorted(out)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_rank_step_equals_single_gpu_step_on_identical_shards():
    port = 29700 + os.getpid() % 200
    two = _run(_worker, 2, (2, port, True))
    one = _run(_single, 1, ())[0]
    for r in two:
        assert abs(r[1] - one[1]) <= 1e-5 * abs(one[1])
        for a, b in zip(r[2:], one[2:]):
            assert abs(a - b) <= 1e-6 * abs(b) + 1e-6


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_ranks_stay_in_lockstep_on_different_shards():
    port = 29900 + os.getpid() % 90
    two = _run(_worker, 2, (2, port, False))
    assert two[0][1] != two[1][1]                    # different data, different losses
    for a, b in zip(two[0][2:], two[1][2:]):
        assert a == b                                # bit-identical parameters on both ranks
