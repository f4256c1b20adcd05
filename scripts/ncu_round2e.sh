set -x
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:cf_seq_factual -s 3 -c 1 -o gpurun_out/r2_k3_factual_v3 python scripts/run_k3.py 1000000 2 > gpurun_out/ncu_k3f.log 2>&1
$NCU -k regex:poly_tsqr_kernel -s 1 -c 1 -o gpurun_out/r2_k4p_tsqr python scripts/run_poly.py 1000000 2 > gpurun_out/ncu_k4p.log 2>&1
