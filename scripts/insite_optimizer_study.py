"""Which optimiser settings reproduce the reference's two INSITE log lines?  
Per-treatment models: results/2_main_table/final_with_insite.txt:2362 (seed 1); joint model: results/ablation/one_ode/...txt:6
(seed 10).  Sweeps gtol / max_iter / the zoom-failure fallback and prints the relative deviation of the 8 RMSEs."""
import json, os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import numpy as np
from b200_insite.config import default_config
from b200_insite.dataset import SyntheticCancerDatasetCollection
from b200_insite.sindy import run_experiment

KEYS = ('encoder_test_rmse_all', 'encoder_test_rmse_orig', 'encoder_test_rmse_last') + tuple(f'decoder_test_rmse_{k}-step' for k in range(2, 7))
gold = lambda name: json.load(open(os.path.join(ROOT, 'tests', 'golden', name)))


def ref_list(log):
    return [log['encoder_test_rmse_all'], log['encoder_test_rmse_orig'], log['encoder_test_rmse_last']] + list(log['decoder_test_rmse_2_to_6_step'])


def main():
    cases = [('per-treatment (seed 1)', dict(seed=1), 'multiclass', ref_list(gold('ref_log_seed1.json')['insite'])),
             ('joint (seed 10)', dict(seed=10, joint_model=True), 'multilabel', ref_list(gold('ref_log_joint_seed10.json')['insite']))]
    for name, kw, mode, ref in cases:
        col = SyntheticCancerDatasetCollection(2.0, 2.0, {'train': 1000, 'val': 100, 'test': 100}, seed=kw['seed'], treatment_mode=mode)
        col.process_data_multi()
        for ls_mode, gtol, mi in [(int(m), g, i) for m in os.environ.get("LS_MODES", "0").split(",")
                                  for g, i in ((1e-5, 3200),)]:
            os.environ["B200I_K7_LS_MODE"] = str(ls_mode)
            for fb in (True, False):
                cfg = default_config(insite=True, treatment_mode=mode, insite_gtol=gtol, insite_max_iter=mi,
                                     insite_zoom_failure_fallback=fb, **kw)
                res, model = run_experiment(cfg, col)
                got = np.array([res[k] for k in KEYS]); r = np.array(ref)
                dev_ = (got - r) / r
                info = model.last_fit_info
                hist = info.get('status_low_byte')
                print(json.dumps({"model": name, "ls_mode": ls_mode, "gtol": gtol, "max_iter": mi, "fallback_on_zoom_failure": fb,
                                  "max_abs_rel_dev": float(np.abs(dev_).max()), "rel_dev": [round(float(v), 5) for v in dev_],
                                  "status_hist_last_eval": None if hist is None else [int(v) for v in hist],
                                  "iterations_mean": info.get('iterations_mean')}), flush=True)


if __name__ == "__main__":
    main()
