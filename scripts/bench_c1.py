"""BASELINE config C1 (the reference's own default case) end to end through the drop-in classes:
SyntheticCancerDatasetCollection (numpy RNG in the reference's order -> K1/K2/K3) -> process_data_multi -> SINDY.fit ->
one-step and tau-step counterfactual RMSEs, for SINDy (population) and INSITE (per-row BFGS), seed 1, gamma 2.
Prints one JSON line per (size, method) with wall-clock seconds per stage and the deviation from the reference's
committed run log (results/2_main_table/final_with_insite.txt:6 / :2362, tests/golden/ref_log_seed1.json), whose own
`seconds_taken` were 13.1 s (SINDy) and 84.8 s (INSITE) on unknown hardware with a cached dataset.
Usage: python scripts/bench_c1.py [--big]     (--big adds the 10k/1k/1k size BASELINE.json quotes; no log to compare)"""
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import numpy as np
import torch
from b200_insite import device as dev
from b200_insite.config import default_config
from b200_insite.dataset import SyntheticCancerDatasetCollection
from b200_insite.sindy import SINDY


def run(sizes, insite, log):
    cfg = default_config(insite=insite, n_train=sizes[0], n_val=sizes[1], n_test=sizes[2])
    t = {}
    t0 = time.perf_counter()
    col = SyntheticCancerDatasetCollection(2.0, 2.0, {'train': sizes[0], 'val': sizes[1], 'test': sizes[2]}, seed=1)
    torch.cuda.synchronize(); t['generate_s'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    col.process_data_multi()
    t['process_s'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    model = SINDY(cfg, col)
    model.fit(col.train_f, col.val_f)
    torch.cuda.synchronize(); t['fit_s'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    orig, all_, last = model.get_normalised_masked_rmse(col.test_cf_one_step, one_step_counterfactual=True)
    torch.cuda.synchronize(); t['one_step_eval_s'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    rm = model.get_normalised_n_step_rmses(col.test_cf_treatment_seq)
    torch.cuda.synchronize(); t['tau_step_eval_s'] = time.perf_counter() - t0
    rows = int(col.test_cf_one_step.data['sequence_lengths'].shape[0] + col.test_cf_treatment_seq.data['sequence_lengths'].shape[0])
    out = {"config": f"C1 {sizes[0]}/{sizes[1]}/{sizes[2]} patients, 60 steps, gamma 2, seed 1", "method": "insite" if insite else "sindy",
           "seconds": {k: round(v, 4) for k, v in t.items()}, "total_s": round(sum(t.values()), 4),
           "test_rows": rows, "rmse_all_orig_last": [all_, orig, last], "rmse_tau_2_to_6": [float(x) for x in rm],
           "equation": model.global_equation_string[:80] + "..."}
    if log is not None:
        ref = [log['encoder_test_rmse_all'], log['encoder_test_rmse_orig'], log['encoder_test_rmse_last']] + list(log['decoder_test_rmse_2_to_6_step'])
        got = [all_, orig, last] + [float(x) for x in rm]
        out["max_rel_dev_vs_reference_log"] = float(np.max(np.abs(np.array(got) - np.array(ref)) / np.abs(ref)))
        out["reference_log_seconds_taken"] = 84.796 if insite else 13.109
        if insite:
            out["individualised_fits_per_s"] = rows / max(t['one_step_eval_s'] + t['tau_step_eval_s'], 1e-9)
    print(json.dumps(out), flush=True)


def main():
    dev.require_cuda()
    log = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_log_seed1.json")))
    run((1000, 100, 100), False, log['sindy'])      # includes the one-off CUDA context / library start-up
    run((1000, 100, 100), False, log['sindy'])
    run((1000, 100, 100), True, log['insite'])
    if "--big" in sys.argv:
        run((10000, 1000, 1000), False, None)
        run((10000, 1000, 1000), True, None)


if __name__ == "__main__":
    main()
