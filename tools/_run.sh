set -x
timeout 500 python bench.py --steps 10 > gpurun_out/r3_bench_final3.log 2>gpurun_out/r3_bench_final3.err; echo rc=$?; tail -c 200 gpurun_out/r3_bench_final3.log; tail -2 gpurun_out/r3_bench_final3.err
