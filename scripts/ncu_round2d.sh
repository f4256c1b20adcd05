set -x
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:cf_one_step_kernel -s 3 -c 1 -o gpurun_out/r2_k2_one_step_v2 python scripts/bench_configs.py 1000000 1000 > gpurun_out/ncu_k2.log 2>&1
$NCU -k regex:cf_seq_factual -s 3 -c 1 -o gpurun_out/r2_k3_factual_v2 python scripts/run_k3.py 1000000 2 > gpurun_out/ncu_k3f.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cf_ -c 40 --csv --log-file gpurun_out/r2_launches_k3.csv python scripts/run_k3.py 1000000 2 > gpurun_out/ncu_k3l.log 2>&1
