set -x
timeout 300 python -m pytest tests/test_conv2d_gpu.py tests/test_depth_engine_gpu.py -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/r3_bench_wgpre.log 2>&1; tail -c 200 gpurun_out/r3_bench_wgpre.log
