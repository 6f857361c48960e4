#!/bin/bash
# Round-2 profile batch (one B200, under gpurun): ncu --set full captures of the dominant kernels and ncu launch lists of
# the three workloads.  Outputs go to gpurun_out/; tools/ncu_summary.py turns the .ncu-rep files into profiles/*.csv.
set -x
ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc2 -s 10 -c 2 -f -o gpurun_out/r02_fwd python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r02_ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc_bwd -s 1 -c 1 -f -o gpurun_out/r02_bwd python tools/tcb_prof.py > gpurun_out/r02_ncu_bwd.log 2>&1
STEPS=1 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 60 -c 20 -f -o gpurun_out/r02_conv python tools/bench_dfnet.py > gpurun_out/r02_ncu_conv.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 700 --csv --log-file gpurun_out/r02_train_launches.csv python tools/bench_train.py --steps 1 --warmup 3 > /dev/null 2>&1
STEPS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_dfnet_launches.csv python tools/bench_dfnet.py > /dev/null 2>&1
ls -la gpurun_out/r02_*
