"""Throughput of the other BASELINE.json configurations on one B200 (CUDA events, median of reps):
  C3  treatment-sequence counterfactual rollouts tau=1..5 (K3) and one-step counterfactuals (K2), compact cohort
  C4  individualised per-patient fits: batched ridge-to-prior STLSQ (K5b), INSITE BFGS (K7), discovered-ODE rollout (K6)
Usage: python scripts/bench_configs.py [N_cf] [N_fit]    -> one JSON line per kernel
"""
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import numpy as np
import torch
from b200_insite import device as dev
from b200_insite import counterfactual as cfm
import b200_insite.cancer_simulation as cs


def timed(fn, reps=3, warm=1):
    r = None
    for _ in range(warm):
        r = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        r = None   # release the previous result first: a 24 GB cohort otherwise forces fresh cudaMallocs inside the timing
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), r


def main():
    n_cf = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    n_fit = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    T, H = 60, 5
    dev.require_cuda()
    g = torch.Generator(device='cuda'); g.manual_seed(7)
    np.random.seed(3)
    params = cs.generate_params(n_cf, 2.0, 2.0, 15, 0)
    block = torch.from_numpy(dev.pack_params(params)).cuda()
    types = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.float64)).cuda()
    # --- C3 / K3: draws per patient randn(T+H), rand(T) x3 (reference order) ---
    noise = 0.01 * torch.randn((n_cf, T + H), generator=g, device='cuda', dtype=torch.float64)
    rec, chemo, radio = (torch.rand((n_cf, T), generator=g, device='cuda', dtype=torch.float64) for _ in range(3))
    ms, coh = timed(lambda: cfm.sim_cf_treatment_seq(block, noise, rec, chemo, radio, T, H))
    steps = float((coh.n_steps.double()).sum().item())
    rows = int(coh.total_rows)
    print(json.dumps({"kernel": "sim_cf_treatment_seq (K3, compact cohort + wavefront)", "patients": n_cf, "ms": ms,
                      "rows": rows, "levels": int(coh.levels), "rows_per_s": rows / ms * 1e3,
                      "rollout_steps_per_s": (steps + 5.0 * rows) / ms * 1e3,
                      "bytes_written": n_cf * ((T - 1) * 2 * H * H * 8 + T * 9 + (T - 1) * 2 + 16),
                      "write_GBs": n_cf * ((T - 1) * 2 * H * H * 8 + T * 9) / ms / 1e6}), flush=True)
    del coh
    # the same from host parameters only: H2D of the parameter block, draws from the device generator, K3
    h_block = torch.from_numpy(dev.pack_params(params)).pin_memory()
    def c3_from_host():
        b = h_block.cuda(non_blocking=True)
        dr = cfm.generated_draws(n_cf, T, H, seed=5)
        return cfm.sim_cf_treatment_seq(b, *dr, T, H)
    ms, coh = timed(c3_from_host)
    print(json.dumps({"kernel": "C3 end to end: pinned host parameters -> H2D -> device-generated draws -> K3", "patients": n_cf,
                      "ms": ms, "rows": int(coh.total_rows), "rows_per_s": int(coh.total_rows) / ms * 1e3,
                      "h2d_bytes": int(h_block.numel() * 8)}), flush=True)
    del coh
    noise1 = noise[:, :T].contiguous()
    del noise
    ms, coh = timed(lambda: cfm.sim_cf_one_step(block, noise1, rec, chemo, radio, T))
    rows = int(coh.total_rows)
    print(json.dumps({"kernel": "sim_cf_one_step (K2, compact cohort + wavefront)", "patients": n_cf, "ms": ms,
                      "rows": rows, "levels": int(coh.levels), "rows_per_s": rows / ms * 1e3}), flush=True)
    del coh, noise1, rec, chemo, radio
    torch.cuda.empty_cache()
    # --- C4: factual cohort -> per-patient fits ---
    n = n_fit
    np.random.seed(4)
    params = cs.generate_params(n, 2.0, 2.0, 15, 0)
    block = torch.from_numpy(dev.pack_params(params)).cuda()
    static = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.float64)).cuda()
    nz = 0.01 * torch.randn((n, T), generator=g, device='cuda', dtype=torch.float64)
    r3 = [torch.rand((n, T), generator=g, device='cuda', dtype=torch.float64) for _ in range(3)]
    out, _ = dev.sim_factual(block, nz, *r3, T)
    del nz, r3
    st = dev.theta_gram(out['cancer_volume'], out['chemo_application'], out['radio_application'], out['sequence_lengths'],
                        static, out['chemo_dosage'], out['radio_dosage'])
    coefs, support = dev.stlsq_population(st)
    codes = dev.treatment_codes(out['chemo_application'], out['radio_application'], T)
    x = out['cancer_volume']
    seq = out['sequence_lengths']
    fit_len = seq.to(torch.int32)
    prior = coefs.contiguous()
    ms, pc = timed(lambda: dev.stlsq_batched(x, codes, fit_len, static, prior, 10.0))
    print(json.dumps({"kernel": "stlsq_batched (K5b, ridge-to-prior STLSQ, FP64)", "rows": n, "ms": ms,
                      "fits_per_s": n / ms * 1e3}), flush=True)
    ms, pred = timed(lambda: dev.ode_rollout(x[:, 0].contiguous(), static, codes[:, :T - 1].contiguous(), pc))
    print(json.dumps({"kernel": "ode_rollout (K6, per-row coefficients, 59 x 5 Euler sub-steps)", "rows": n, "ms": ms,
                      "rows_per_s": n / ms * 1e3, "euler_steps_per_s": n * 59 * 5 / ms * 1e3}), flush=True)
    ms32, pred32 = timed(lambda: dev.ode_rollout(x[:, 0].contiguous(), static, codes[:, :T - 1].contiguous(), pc, fp32=True))
    dev_rel = float(((pred32 - pred).abs() / pred.abs().clamp_min(1e-3 * float(pred.abs().max()))).max().item())
    print(json.dumps({"kernel": "ode_rollout_f32 (K6 in float32)", "rows": n, "ms": ms32, "rows_per_s": n / ms32 * 1e3,
                      "max_rel_dev_vs_fp64": dev_rel}), flush=True)
    nb = min(n, int(os.environ.get("BFGS_ROWS", "200000")))
    ms, (c7, status, fval) = timed(lambda: dev.insite_bfgs(x[:nb].contiguous(), codes[:nb].contiguous(),
                                                           fit_len[:nb].contiguous(), 1, static[:nb].contiguous(), prior, 10.0),
                                   reps=2)
    print(json.dumps({"kernel": "insite_bfgs (K7, 16 coefficients per row, FP64)", "rows": nb, "ms": ms,
                      "fits_per_s": nb / ms * 1e3,
                      "status_hist (0 converged, 1 max_iter, 3/5 line search exhausted, 4 perfect start, 6 kept theta0)":
                          np.bincount(status.cpu().numpy().astype(np.int64) & 255, minlength=7).tolist(),
                      "mean_iterations": float((status.cpu().numpy().astype(np.int64) >> 8).mean()),
                      "mean_objective_ratio": float((fval[:, 1] / fval[:, 0].clamp_min(1e-300)).mean().item())}), flush=True)


if __name__ == "__main__":
    main()
