set -x
timeout 500 python -m pytest tests/test_depth_engine_gpu.py tests/test_step_gpu.py tests/test_step_benchconfig_gpu.py tests/test_reproject_gpu.py tests/test_eval_path_gpu.py tests/test_checkpoint_compat_gpu.py tests/test_conv2d_gpu.py -x -q 2>&1 | tail -4
timeout 200 python bench.py --pairs 1 --steps 30 --no-extras --no-cpu-baseline > gpurun_out/r3_bench_lanes_B1.log 2>&1; tail -c 250 gpurun_out/r3_bench_lanes_B1.log
DVD_LANES=1 timeout 200 python bench.py --pairs 1 --steps 30 --no-extras --no-cpu-baseline > gpurun_out/r3_bench_nolanes_B1.log 2>&1; tail -c 250 gpurun_out/r3_bench_nolanes_B1.log
timeout 200 python bench.py --pairs 2 --steps 20 --no-extras --no-cpu-baseline > gpurun_out/r3_bench_lanes_B2.log 2>&1; tail -c 250 gpurun_out/r3_bench_lanes_B2.log
DVD_LANES=1 timeout 200 python bench.py --pairs 2 --steps 20 --no-extras --no-cpu-baseline > gpurun_out/r3_bench_nolanes_B2.log 2>&1; tail -c 250 gpurun_out/r3_bench_nolanes_B2.log
timeout 100 python tools/bench_reproject.py 64 2>&1 | tail -4
