"""Evaluation of the discovered ODE on compact counterfactual cohorts (BASELINE config C3).

The reference scores one dense row per (patient, t, option) -- 227 one-step and 562 sequence rows per test patient,
0.9 TB at 1M patients: ``SINDY.get_predictions`` / ``get_autoregressive_predictions`` (sindy.py:371-431, 433-715,
717-760) and the RMSE methods of TimeVaryingCausalModel (time_varying_model.py:236-313).  Here the cohort stays in the
per-patient form the generators write (counterfactual.CompactCohort) and three kernels do the work:

    K7 / K5b per (patient, t)   one individualisation fit per (patient, t) instead of one per row (the 4 one-step rows
                                and the <= 2H sequence rows of a (patient, t) repeat the same fit problem)
    K8 cf_eval_one_step         rollout + masked squared-error sums of the one-step cohort
    K9 cf_eval_treatment_seq    factual-prefix rollout + H projected steps per valid option, per-horizon sums

Only the error sums leave the kernels; with a process group they are all-reduced (they are additive over patients).
"""
import numpy as np
import torch

from . import device as dev
from .cohort import allreduce_stats


def _finish(sums):
    allreduce_stats(sums)
    torch.cuda.current_stream().synchronize()
    return sums.cpu().numpy()


def one_step_rmses(sums, W, norm_const, percentage=True, scale=1.0):
    """(rmse_orig, rmse_all, rmse_last) from the (3W+2,) sums -- get_normalised_masked_rmse, time_varying_model.py:247-281.
    scale: divide the errors by it (the output std when exp.unscale_rmse is False)."""
    s = np.asarray(sums, dtype=np.float64)
    se, cnt = s[:W] / scale ** 2, s[W:2 * W]
    with np.errstate(invalid='ignore', divide='ignore'):
        mse_orig = (se / cnt).mean()                     # a column without active rows gives NaN, as in the reference
    k = 100.0 if percentage else 1.0
    rmse_orig = np.sqrt(mse_orig) / norm_const * k
    rmse_all = np.sqrt(se.sum() / cnt.sum()) / norm_const * k
    rmse_last = np.sqrt(s[3 * W] / scale ** 2 / s[3 * W + 1]) / norm_const * k
    return rmse_orig, rmse_all, rmse_last


def n_step_rmses(sums, H, norm_const, percentage=True, scale=1.0):
    """Per-horizon RMSEs from the (2H,) sums -- get_normalised_n_step_rmses, time_varying_model.py:298-311."""
    s = np.asarray(sums, dtype=np.float64)
    with np.errstate(invalid='ignore', divide='ignore'):
        r = np.sqrt(s[:H] / scale ** 2 / s[H:2 * H]) / norm_const
    return r * (100.0 if percentage else 1.0)


def individualise(cohort, static, theta0, estimator='bfgs_rollout', lam=10.0, ridge_prior_lam=1e4, threshold=1e-3,
                  zoom_failure_fallback=None, dt=dev.STANDARD_DT, gtol=None, max_iter=None, line_search='jax'):
    """Per-(patient, t) coefficient matrices (n, T-1, 4, 4) of a compact cohort + diagnostics.
    The fit window of the one-step rows is their first t transitions (projection_horizon 1, sindy.py:433), of the
    sequence rows their first t+1 (projection_horizon H, :747)."""
    fit_offset = 0 if cohort.kind == 'one_step' else 1
    if zoom_failure_fallback is None:      # the reference's rule (sindy.py:628-631) belongs to jax's failure semantics
        zoom_failure_fallback = line_search == 'jax'
    if estimator == 'bfgs_rollout':
        coefs, status, fval = dev.insite_bfgs_prefix(cohort.factual, cohort.codes, cohort.n_steps, static, theta0, lam,
                                                     fit_offset, gtol=gtol, max_iter=max_iter, dt=dt, line_search=line_search)
        if zoom_failure_fallback:     # sindy.py:628-631
            failed = (status & 255) == 3
            coefs[failed] = theta0.reshape(4, 4)
        return coefs, {'status': status, 'fval': fval}
    if estimator == 'ridge_prior_stlsq':
        coefs = dev.stlsq_prefix(cohort.factual, cohort.codes, cohort.n_steps, static, theta0, ridge_prior_lam, fit_offset,
                                 threshold=threshold, fd_dt=dt)
        return coefs, {}
    raise ValueError(f"unknown individualisation estimator {estimator!r}")


def evaluate(cohort, static, coefs, drop_below, dt=dev.STANDARD_DT, substeps=dev.STEPS_FOR_DT):
    """Error sums of one compact cohort on this rank (device tensor; see device.cf_eval_*)."""
    if cohort.kind == 'one_step':
        return dev.cf_eval_one_step(cohort.factual, cohort.codes, cohort.cf, cohort.n_steps, static, coefs,
                                    drop_below=drop_below, dt=dt, substeps=substeps)
    return dev.cf_eval_treatment_seq(cohort.factual, cohort.codes, cohort.cf, cohort.valid, cohort.n_steps, static, coefs,
                                     drop_below=drop_below, dt=dt, substeps=substeps)


def evaluate_model(one_step, static_one, seq, static_seq, population_coefs, insite=False, estimator='bfgs_rollout',
                   lam=10.0, ridge_prior_lam=1e4, threshold=1e-3, zoom_failure_fallback=None,
                   norm_const=dev.TUMOUR_DEATH_THRESHOLD, percentage=True, scale=1.0, dt=dev.STANDARD_DT, info=None):
    """The eight test metrics of train_sindy.main (:72-112) from two compact cohorts: encoder_test_rmse_{all,orig,last}
    and decoder_test_rmse_{2..H+1}-step.  population_coefs: (4,4) device tensor (SINDY.joint_coefs)."""
    out = {}
    for cohort, static in ((one_step, static_one), (seq, static_seq)):
        if cohort is None:
            continue
        if insite:
            coefs, diag = individualise(cohort, static, population_coefs, estimator, lam, ridge_prior_lam, threshold,
                                        zoom_failure_fallback, dt)
            if info is not None:
                info[cohort.kind] = diag
            sums = _finish(evaluate(cohort, static, coefs, -1.0, dt))
        else:
            sums = _finish(evaluate(cohort, static, population_coefs, 1e-3, dt))
        if cohort.kind == 'one_step':
            orig, all_, last = one_step_rmses(sums, cohort.T - 1, norm_const, percentage, scale)
            out.update({'encoder_test_rmse_all': all_, 'encoder_test_rmse_orig': orig, 'encoder_test_rmse_last': last})
        else:
            r = n_step_rmses(sums, cohort.H, norm_const, percentage, scale)
            out.update({f'decoder_test_rmse_{k + 2}-step': v for k, v in enumerate(r)})
    return out
