set -x
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r3_tests5.log 2>&1; tail -6 gpurun_out/r3_tests5.log
timeout 500 python bench.py > gpurun_out/r3_bench_full.log 2>gpurun_out/r3_bench_full.err; tail -c 1500 gpurun_out/r3_bench_full.log; tail -3 gpurun_out/r3_bench_full.err
timeout 200 python bench.py --height 448 --width 768 --pairs 4 --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r3_bench_768x448.log 2>&1; tail -c 400 gpurun_out/r3_bench_768x448.log
timeout 200 python bench.py --height 288 --width 512 --frames 200 --pairs 8 --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r3_bench_512x288.log 2>&1; tail -c 400 gpurun_out/r3_bench_512x288.log
