set -x
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/r3_tests7.log 2>&1; tail -4 gpurun_out/r3_tests7.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 500 python bench.py > gpurun_out/r3_bench_final2.log 2>gpurun_out/r3_bench_final2.err; tail -c 200 gpurun_out/r3_bench_final2.log
