set -x
DFB_TC_CTA_GROUP=1 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2r_bench_cg1.json 2> gpurun_out/r2r.err
ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc2 -s 10 -c 2 -f -o gpurun_out/r2r_fwd python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2r_ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc_bwd -s 2 -c 1 -f -o gpurun_out/r2r_bwd python tools/tcb_prof.py > gpurun_out/r2r_ncu_bwd.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 500 --csv --log-file gpurun_out/r2r_train_launches.csv python tools/bench_train.py --steps 1 --warmup 3 > /dev/null 2>&1
STEPS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2r_dfnet_launches.csv python tools/bench_dfnet.py > /dev/null 2>&1
python -c "
import json; d=json.load(open('gpurun_out/r2r_bench_cg1.json')); print('cg1', d['value'], d['roofline']['frac'])"
ls -la gpurun_out/r2r*
