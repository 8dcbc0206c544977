set -x
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r3_tests6.log 2>&1; tail -4 gpurun_out/r3_tests6.log
timeout 500 python bench.py > gpurun_out/r3_bench_final.log 2>gpurun_out/r3_bench_final.err; tail -c 300 gpurun_out/r3_bench_final.log
