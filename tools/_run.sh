python -m pytest tests -m gpu -q > gpurun_out/r3_tests3.log 2>&1; tail -6 gpurun_out/r3_tests3.log
python tools/profile_convs.py r3d > gpurun_out/r3d_conv_table.log 2>&1; head -4 gpurun_out/r3d_conv_table.log
python bench.py --no-extras --no-cpu-baseline > gpurun_out/r3_bench_4.log 2>&1; tail -c 500 gpurun_out/r3_bench_4.log
