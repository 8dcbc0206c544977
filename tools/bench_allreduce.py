"""Time the flat gradient all-reduce (422.6 MB fp32 = MiDaS + MLP) over NCCL: torchrun --nproc-per-node N tools/bench_allreduce.py"""
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    n = 105_660_932
    buf = torch.ones(n, device='cuda')
    for _ in range(3):
        dist.all_reduce(buf)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dist.all_reduce(buf)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    if rank == 0:
        t = ts[len(ts) // 2] * 1e-3
        print({'world': world, 'bytes': 4 * n, 'ms': t * 1e3, 'algbw_GBps': 4 * n / t / 1e9,
               'busbw_GBps': 2 * (world - 1) / world * 4 * n / t / 1e9})
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
