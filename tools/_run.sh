set -x
timeout 300 python -m pytest tests/test_conv2d_gpu.py tests/test_depth_engine_gpu.py -x -q 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r3_bench_2gpu_full.log 2>gpurun_out/r3_bench_2gpu_full.err; echo "rc=$?"; tail -c 600 gpurun_out/r3_bench_2gpu_full.log; tail -5 gpurun_out/r3_bench_2gpu_full.err
