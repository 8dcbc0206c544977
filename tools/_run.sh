set -x
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r3_bench_8gpu.log 2>gpurun_out/r3_bench_8gpu.err; echo "rc=$?"; tail -c 500 gpurun_out/r3_bench_8gpu.log; tail -3 gpurun_out/r3_bench_8gpu.err
