"""Times the compact-cohort evaluation (K8 / K9 and the per-(patient, t) fits) on generated cohorts (profiling target).
Usage: python scripts/run_cf_eval.py N_seq N_one [N_bfgs]"""
import json, os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import numpy as np, torch
from b200_insite import device as dev
from b200_insite import counterfactual as cfm
from b200_insite import compact_eval as ce
import b200_insite.cancer_simulation as cs

n_seq, n_one = int(sys.argv[1]), int(sys.argv[2])
n_bfgs = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
T, H = 60, 5
dev.require_cuda()
COEFS = [[-0.05601456082026624, -0.11598756834077, -0.07958279124512227, 0.07326347275734027],
         [-0.5517350343589641, -0.8761667536689084, -0.053397817822270766, -0.035996455669168224],
         [-3.649800303098579, -0.8626472911638889, 1.1157911611997717, -0.6373790072514276],
         [-1.6336216074116419, -3.49858670473956, -3.584018882175004, 0.06151172618047967]]
coefs = dev.to_device(np.array(COEFS))


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), r


def cohort(kind, n, seed):
    np.random.seed(seed)
    params = cs.generate_params(n, 2.0, 2.0, 15, 0)
    block = torch.from_numpy(dev.pack_params(params)).cuda()
    static = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.float64)).cuda()
    dr = cfm.generated_draws(n, T, H if kind == 'seq' else 0, seed=seed)
    c = cfm.sim_cf_treatment_seq(block, *dr, T, H) if kind == 'seq' else cfm.sim_cf_one_step(block, *dr, T)
    return c, static


for kind, n in (('seq', n_seq), ('one', n_one)):
    c, static = cohort(kind, n, 5)
    ms, sums = timed(lambda: ce.evaluate(c, static, coefs, 1e-3))
    bytes_read = n * ((T - 1) * 2 * H * H * 8 + T * 9 + (T - 1) * 2 + 12) if kind == 'seq' else n * ((T - 1) * 4 * 8 + T * 9 + 12)
    s = sums.cpu().numpy()
    rm = ce.n_step_rmses(s, H, dev.TUMOUR_DEATH_THRESHOLD) if kind == 'seq' else ce.one_step_rmses(s, T - 1, dev.TUMOUR_DEATH_THRESHOLD)
    print(json.dumps({"kernel": f"cf_eval_{kind} (population coefficients)", "patients": n, "rows": int(c.total_rows), "ms": ms,
                      "read_GBs": bytes_read / ms / 1e6, "rows_per_s": c.total_rows / ms * 1e3,
                      "rmse": [float(v) for v in np.atleast_1d(rm)]}), flush=True)
    per = coefs.reshape(1, 1, 4, 4).expand(n, T - 1, 4, 4).contiguous()
    ms, sums2 = timed(lambda: ce.evaluate(c, static, per, 1e-3))
    print(json.dumps({"kernel": f"cf_eval_{kind} (per-(patient,t) coefficients)", "patients": n, "ms": ms,
                      "read_GBs": (bytes_read + n * (T - 1) * 128) / ms / 1e6,
                      "max_rel_vs_shared": float(((sums2 - sums).abs() / sums.abs().clamp_min(1e-300)).max().item())}), flush=True)
    ms, pc = timed(lambda: dev.stlsq_prefix(c.factual, c.codes, c.n_steps, static, coefs, 1e4, 0 if kind == 'one' else 1), reps=3)
    print(json.dumps({"kernel": f"stlsq_prefix ({kind})", "patients": n, "fits": n * (T - 1), "ms": ms,
                      "fits_per_s": n * (T - 1) / ms * 1e3}), flush=True)
    del per, pc
    nb = min(n, n_bfgs)
    sub = lambda a: a[:nb].contiguous()
    ms, (bc, st, fv) = timed(lambda: dev.insite_bfgs_prefix(sub(c.factual), sub(c.codes), sub(c.n_steps), sub(static), coefs,
                                                            10.0, 0 if kind == 'one' else 1), reps=2)
    stn = st.cpu().numpy().astype(np.int64)
    done = stn >= 0
    print(json.dumps({"kernel": f"insite_bfgs_prefix ({kind})", "patients": nb, "fits": int(done.sum()), "ms": ms,
                      "fits_per_s": float(done.sum()) / ms * 1e3,
                      "status_hist": np.bincount(stn[done] & 255, minlength=7).tolist(),
                      "mean_iterations": float((stn[done] >> 8).mean())}), flush=True)
    del c, bc, st, fv
    torch.cuda.empty_cache()
