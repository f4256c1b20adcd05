set -x
NCU="ncu --set full --clock-control none --import-source on"
python -m pytest tests/test_gpu_compact_eval.py tests/test_gpu_irregular.py tests/test_gpu_insite.py tests/test_gpu_fit.py -x -q -m gpu > gpurun_out/r2_c9_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_c9_pytest.log
python scripts/run_c4.py 1000000 3 > gpurun_out/r2_c4_plain2.log 2>&1
python scripts/run_c4.py 1000000 3 irregular >> gpurun_out/r2_c4_plain2.log 2>&1
python scripts/run_cf_eval.py 1000000 1000000 1000 > gpurun_out/r2_cfeval2.log 2>&1
$NCU -k regex:ode_rollout_tiled -s 2 -c 1 -o gpurun_out/r2_k6_tiled_v2 python scripts/run_c4.py 1000000 2 > gpurun_out/ncu_k6.log 2>&1
$NCU -k regex:stlsq_batched_tiled -s 2 -c 1 -o gpurun_out/r2_k5b_tiled_v2 python scripts/run_c4.py 1000000 2 > gpurun_out/ncu_k5b.log 2>&1
$NCU -k regex:cf_eval_seq -s 1 -c 1 -o gpurun_out/r2_k9_eval_seq_v2 python scripts/run_cf_eval.py 1000000 1000 1000 > gpurun_out/ncu_k9.log 2>&1
$NCU -k regex:cf_eval_one_step -s 1 -c 1 -o gpurun_out/r2_k8_eval_one_v2 python scripts/run_cf_eval.py 1000 1000000 1000 > gpurun_out/ncu_k8.log 2>&1
$NCU -k regex:insite_bfgs_kernel -c 1 -o gpurun_out/r2_k7_bfgs python scripts/run_k7.py > gpurun_out/ncu_k7.log 2>&1
tail -3 gpurun_out/r2_c9_pytest.log; cat gpurun_out/r2_c4_plain2.log; cat gpurun_out/r2_cfeval2.log
