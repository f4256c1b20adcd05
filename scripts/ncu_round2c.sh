set -x
NCU="ncu --set full --clock-control none --import-source on"
python scripts/run_cf_eval.py 1000000 1000000 1000 > gpurun_out/r2_cfeval3.log 2>&1
python scripts/run_c4.py 1000000 3 > gpurun_out/r2_c4_plain3.log 2>&1
python scripts/run_k7.py 200000 3 > gpurun_out/r2_k7_plain.log 2>&1
$NCU -k regex:cf_eval_one_step -s 1 -c 1 -o gpurun_out/r2_k8_eval_one_v3 python scripts/run_cf_eval.py 1000 1000000 1000 > gpurun_out/ncu_k8.log 2>&1
$NCU -k regex:insite_bfgs_kernel -c 1 -o gpurun_out/r2_k7_bfgs_jax python scripts/run_k7.py 200000 1 > gpurun_out/ncu_k7.log 2>&1
$NCU -k regex:cf_one_step_kernel -s 3 -c 1 -o gpurun_out/r2_k2_one_step python scripts/bench_configs.py 1000000 1000 > gpurun_out/ncu_k2.log 2>&1
$NCU -k regex:cf_seq_project -s 1 -c 1 -o gpurun_out/r2_k3_project python scripts/run_k3.py 1000000 2 > gpurun_out/ncu_k3p.log 2>&1
$NCU -k regex:cf_seq_factual -s 3 -c 1 -o gpurun_out/r2_k3_factual python scripts/run_k3.py 1000000 2 > gpurun_out/ncu_k3f.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --insite-patients 2000 > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/r2_cfeval3.log gpurun_out/r2_c4_plain3.log gpurun_out/r2_k7_plain.log
