set -x
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r3_launches_step.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r3_ncu_bench.log 2>&1; tail -c 300 gpurun_out/r3_ncu_bench.log; wc -l gpurun_out/r3_launches_step.csv
timeout 200 $NCU -k regex:conv2d_tc -s 3 -c 1 -o gpurun_out/r3_ncu_conv_fwd_3x3_256 -f python tools/bench_conv_layer.py fwd 16 56 96 256 256 3 1 1 2 > gpurun_out/r3_ncu_a.log 2>&1; tail -2 gpurun_out/r3_ncu_a.log
timeout 200 $NCU -k regex:conv2d_tc -s 3 -c 1 -o gpurun_out/r3_ncu_conv_fwd_1x1_1024 -f python tools/bench_conv_layer.py fwd 16 14 24 1024 1024 1 1 1 2 > gpurun_out/r3_ncu_b.log 2>&1; tail -2 gpurun_out/r3_ncu_b.log
timeout 200 $NCU -k regex:conv_wgrad -s 3 -c 1 -o gpurun_out/r3_ncu_conv_wgrad_1x1_1024 -f python tools/bench_conv_layer.py wgrad 16 14 24 1024 1024 1 1 1 2 > gpurun_out/r3_ncu_c.log 2>&1; tail -2 gpurun_out/r3_ncu_c.log
timeout 200 $NCU -k regex:conv2d_tc -s 3 -c 1 -o gpurun_out/r3_ncu_conv_fwd_grouped_1024 -f python tools/bench_conv_layer.py fwd 16 14 24 1024 1024 3 1 32 2 > gpurun_out/r3_ncu_d.log 2>&1; tail -2 gpurun_out/r3_ncu_d.log
timeout 300 $NCU -k regex:mlp_ -s 12 -c 8 -o gpurun_out/r3_ncu_mlp -f python tools/bench_mlp.py 2 > gpurun_out/r3_ncu_e.log 2>&1; tail -3 gpurun_out/r3_ncu_e.log
ls -la gpurun_out/*.ncu-rep | tail -8
