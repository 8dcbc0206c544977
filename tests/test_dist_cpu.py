"""CPU, world_size 2, gloo: host-side logic of the data-parallel path — flat parameter
broadcast, flat gradient all-reduce (sum; the 1/world factor goes into the Adam kernel), epoch-metric
reduction, and the pair partition across ranks. No CUDA kernels involved."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from dvd_b200.flat import FlatParams
    from dvd_b200.models.netinterface import NetInterface
    from dvd_b200.networks.sceneflow_field import SceneFlowFieldNet
    torch.manual_seed(100 + rank)                       # ranks start from DIFFERENT weights
    net = SceneFlowFieldNet(net_width=256, n_layers=4, time_dependent=True, N_freq_xyz=16, N_freq_t=16)
    flat = FlatParams(net)
    names = [n for n, _ in net.named_parameters()]
    # 1) one flat broadcast makes every rank identical to rank 0, through the per-tensor views
    flat.broadcast(0)
    ref = [torch.zeros_like(flat.data) for _ in range(world)]
    dist.all_gather(ref, flat.data)
    same = all(torch.equal(ref[0], r) for r in ref)
    view_ok = torch.equal(dict(net.named_parameters())['convs.3.conv.weight'].reshape(-1),
                          flat.data[flat.offsets[names.index('convs.3.conv.weight')]:][:256 * 256])
    # 2) gradient all-reduce = sum over ranks, visible through p.grad views
    flat.zero_grad()
    for p in net.parameters():
        p.grad.add_(float(rank + 1))
    flat.allreduce_grad()
    g_ok = all(bool((p.grad == 3.0).all()) for p in net.parameters())   # 1 + 2
    # 3) epoch metrics: mean over ranks in one all-reduce
    shell = NetInterface.__new__(NetInterface)
    shell.device = torch.device('cpu')
    red = shell._reduce_epoch_log({'loss': float(rank), 'acc_reg': 2.0})
    m_ok = abs(red['loss'] - 0.5) < 1e-12 and abs(red['acc_reg'] - 2.0) < 1e-12
    q.put((rank, same, view_ok, g_ok, m_ok))
    dist.destroy_process_group()


def test_flat_broadcast_allreduce_and_metric_reduce_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, view_ok, g_ok, m_ok in res:
        assert same and view_ok and g_ok and m_ok, (rank, same, view_ok, g_ok, m_ok)


def test_pair_partition_is_disjoint_and_gap_uniform():
    sys.path.insert(0, ROOT)
    import bench
    for step in range(10):
        seen = set()
        gaps = set()
        for rank in range(8):
            gap, pairs = bench.step_pairs(step, rank, 8, 4)
            gaps.add(gap)
            for p in pairs:
                assert p[1] - p[0] == gap and 0 <= p[0] and p[1] < bench.N_FRAMES
                seen.add(p)
        assert len(gaps) == 1                 # every rank runs the same number of Euler steps in a step
        assert len(seen) == 32                # 8 ranks x 4 pairs: disjoint


def test_gap_bucket_sampler_partitions_the_pair_list():
    """product sampler (datasets/resident.py): per epoch every rank gets gap-uniform steps, the ranks' pairs of one step are
    disjoint, nothing is drawn twice in an epoch, and epochs reshuffle."""
    from dvd_b200 import synthetic
    from dvd_b200.datasets.resident import GapBucketSampler
    pairs = synthetic.pair_list(80, (1, 2, 4, 6, 8))
    gaps = [b - a for a, b in pairs]
    world, B = 8, 4
    samplers = [GapBucketSampler(gaps, B, world, r, seed=3) for r in range(world)]
    assert len({len(s) for s in samplers}) == 1 and len(samplers[0]) == sum((gaps.count(g) // (world * B)) for g in set(gaps))
    per_rank = [list(iter(s)) for s in samplers]
    drawn = set()
    for step in range(len(per_rank[0])):
        step_gaps = set()
        for r in range(world):
            idx = per_rank[r][step]
            assert len(idx) == B
            for i in idx:
                assert i not in drawn
                drawn.add(i)
                step_gaps.add(gaps[i])
        assert len(step_gaps) == 1
    e0 = list(iter(samplers[0]))
    samplers[0].set_epoch(1)
    assert list(iter(samplers[0])) != e0
    samplers[0].set_epoch(0)
    assert list(iter(samplers[0])) == e0


def test_resident_sequence_rebuilds_the_pair_file_batch():
    """ResidentSequence on the CPU device (pure indexing): the assembled batch equals the collated pair files."""
    import torch
    from dvd_b200.datasets import get_dataset
    from dvd_b200.datasets.resident import ResidentSequence
    from dvd_b200.options import options_train
    opt, _ = options_train.parse(['--net', 'scene_flow_motion_field', '--dataset', 'synthetic_sequence', '--gaps', '2,4',
                                  '--height', '16', '--width', '24', '--n_frames', '12'])
    ds = get_dataset('synthetic_sequence')(opt)
    seq = ResidentSequence(ds, 'cpu')
    assert len(seq) == len(ds) and seq.images.shape[0] <= 12
    pick = [i for i, g in enumerate(seq.gaps) if g == 4][:3]
    b = seq.batch(pick)
    assert b['steps_hint'] == 4 and b['img_1'].shape == (3, 3, 16, 24)
    for j, i in enumerate(pick):
        it = ds[i]
        for k in ('flow_1_2', 'mask_2', 'R_2_T', 't_1', 'K_inv', 'time_stamp_2'):
            assert torch.equal(b[k][j], it[k][0]), k
        # frames are de-duplicated by frame id: the synthetic generator draws img_2 per PAIR, the cache keeps the first image seen
        assert b['img_1'][j].shape == it['img_1'][0].shape
