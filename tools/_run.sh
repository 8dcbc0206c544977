set -x
timeout 120 python tools/probe_halo.py > gpurun_out/r3_probe_halo.log 2>&1; cat gpurun_out/r3_probe_halo.log
timeout 400 python -m pytest tests/test_conv2d_gpu.py tests/test_mlp_gpu.py tests/test_depth_engine_gpu.py tests/test_step_gpu.py tests/test_step_benchconfig_gpu.py -x -q > gpurun_out/r3_tests4.log 2>&1; tail -6 gpurun_out/r3_tests4.log
timeout 200 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/r3_bench_ov1.log 2>&1; tail -c 400 gpurun_out/r3_bench_ov1.log
DVD_BWD_OVERLAP=0 timeout 200 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/r3_bench_ov0.log 2>&1; tail -c 400 gpurun_out/r3_bench_ov0.log
DVD_CONV_NT=0 timeout 200 python bench.py --steps 20 --no-extras --no-cpu-baseline > gpurun_out/r3_bench_ov1_nt0.log 2>&1; tail -c 400 gpurun_out/r3_bench_ov1_nt0.log
