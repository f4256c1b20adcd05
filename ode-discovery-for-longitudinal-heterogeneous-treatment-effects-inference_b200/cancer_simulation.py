"""Drop-in replacement for the reference module ``src.data.cancer_sim.cancer_simulation``
(libs_m/ct/src/data/cancer_sim/cancer_simulation.py): same function names, arguments, returned
dict keys, shapes and dtypes, same consumption of the global ``np.random``
stream -- but the per-patient time stepping runs in hand-written CUDA (csrc/sim_*.cu) on a B200.

    generate_params(num_patients, chemo_coeff, radio_coeff, window_size, lag)      <- :66-93 (+ :96-215)
    simulate_factual(simulation_params, seq_length, assigned_actions=None)         <- :218-375
    simulate_counterfactual_1_step(simulation_params, seq_length)                  <- :378-563
    simulate_counterfactuals_treatment_seq(simulation_params, seq_length,
                                           projection_horizon, cf_seq_mode)        <- :566-773
    get_scaling_params(sim)                                                        <- :776-796

Parameter generation stays on the host: it is O(N), uses scipy's truncnorm and consumes the RNG
data-dependently (rejection sampling), so it has to be the same numpy/scipy calls in the same order.
"""
import logging

import numpy as np
import pandas as pd
from scipy.stats import truncnorm

logger = logging.getLogger(__name__)


def calc_volume(diameter):
    return 4 / 3 * np.pi * (diameter / 2) ** 3


def calc_diameter(volume):
    return ((volume / (4 / 3 * np.pi)) ** (1 / 3)) * 2


TUMOUR_CELL_DENSITY = 5.8 * 10 ** 8
TUMOUR_DEATH_THRESHOLD = calc_volume(13)

# (mu, sigma, lower bound, upper bound) of the log-normal initial diameter per stage
tumour_size_distributions = {'I': (1.72, 4.70, 0.3, 5.0),
                             'II': (1.96, 1.63, 0.3, 13.0),
                             'IIIA': (1.91, 9.40, 0.3, 13.0),
                             'IIIB': (2.76, 6.87, 0.3, 13.0),
                             'IV': (3.86, 8.82, 0.3, 13.0)}
cancer_stage_observations = {'I': 1432, 'II': 128, 'IIIA': 1306, 'IIIB': 7248, 'IV': 12840}


def get_standard_params(num_patients):
    """Static per-patient parameters; RNG draw order of the reference (:112, :135, :165, :177, :190, :209)."""
    n = int(num_patients)
    stages = sorted(tumour_size_distributions)
    total_obs = sum(cancer_stage_observations.values())
    stage_draw = np.random.choice(stages, n, p=[cancer_stage_observations[s] / total_obs for s in stages])

    diameters, stage_labels = [], []
    for stage in stages:
        count = int(np.count_nonzero(stage_draw == stage))
        mu, sigma, lo, hi = tumour_size_distributions[stage]
        z = truncnorm.rvs((np.log(lo) - mu) / sigma, (np.log(hi) - mu) / sigma, size=count)
        diameters.append(np.exp((z * sigma) + mu))
        stage_labels.append(np.full(count, stage, dtype='<U4'))
    diameters = np.concatenate(diameters) if diameters else np.zeros(0)
    stage_labels = np.concatenate(stage_labels) if stage_labels else np.zeros(0, dtype='<U4')

    rho_mean, rho_sd = 7 * 10 ** -5, 7.23 * 10 ** -3
    alpha_mean, alpha_sd = 0.0398, 0.168
    beta_c_mean, beta_c_sd = 0.028, 0.0007
    cross = 0.87 * alpha_sd * rho_sd
    cov = np.array([[alpha_sd ** 2, cross], [cross, rho_sd ** 2]])
    mean = np.array([alpha_mean, rho_mean])
    kept, n_kept = [], 0
    while n_kept < n:       # rejection sampling: both components must be positive
        draw = np.random.multivariate_normal(mean, cov, size=n)
        ok = draw[(draw[:, 0] > 0.0) & (draw[:, 1] > 0.0)]
        kept.append(ok)
        n_kept += ok.shape[0]
    alpha_rho = np.concatenate(kept, axis=0)[:n] if kept else np.zeros((0, 2))

    patient_types = np.random.choice([1, 2, 3], n)
    chemo_adj = np.where(patient_types < 3, 0.0, 0.1)
    radio_adj = np.where(patient_types > 1, 0.0, 0.1)
    alpha = alpha_rho[:, 0] + alpha_mean * radio_adj
    rho = alpha_rho[:, 1]
    beta = alpha / 10
    beta_c = beta_c_mean + beta_c_sd * truncnorm.rvs((0.0 - beta_c_mean) / beta_c_sd,
                                                     (np.inf - beta_c_mean) / beta_c_sd,
                                                     size=n) + beta_c_mean * chemo_adj
    holder = {'patient_types': patient_types,
              'initial_stages': stage_labels,
              'initial_volumes': calc_volume(diameters),
              'alpha': alpha, 'rho': rho, 'beta': beta, 'beta_c': beta_c,
              'K': np.full(n, calc_volume(30))}
    order = list(range(n))
    np.random.shuffle(order)     # python list, as the reference shuffles one
    return {k: v[order] for k, v in holder.items()}


def generate_params(num_patients, chemo_coeff, radio_coeff, window_size, lag):
    params = get_standard_params(num_patients)
    n = params['patient_types'].shape[0]
    d_max = calc_diameter(TUMOUR_DEATH_THRESHOLD)
    params['chemo_sigmoid_intercepts'] = np.full(n, d_max / 2.0)
    params['radio_sigmoid_intercepts'] = np.full(n, d_max / 2.0)
    params['chemo_sigmoid_betas'] = np.full(n, chemo_coeff / d_max)
    params['radio_sigmoid_betas'] = np.full(n, radio_coeff / d_max)
    params['window_size'] = window_size
    params['lag'] = lag
    return params


# --------------------------------------------------------------------------------------------------
# simulators
# --------------------------------------------------------------------------------------------------
def _draw_per_patient(n, seq_length, extra):
    """Per-patient draw order of the two counterfactual generators (:440-453, :640-653)."""
    noise = np.empty((n, seq_length + extra))
    rec = np.empty((n, seq_length))
    chemo = np.empty((n, seq_length))
    radio = np.empty((n, seq_length))
    for i in range(n):
        noise[i] = 0.01 * np.random.randn(seq_length + extra)
        rec[i] = np.random.rand(seq_length)
        chemo[i] = np.random.rand(seq_length)
        radio[i] = np.random.rand(seq_length)
    return noise, rec, chemo, radio


def simulate_factual(simulation_params, seq_length, assigned_actions=None):
    """Factual trajectories (train / validation subsets).  Reference: :218-375."""
    import torch
    from . import device as dev
    dev.require_cuda()
    n = simulation_params['initial_stages'].shape[0]
    # bulk draws in the reference's order (:275-279)
    noise = 0.01 * np.random.randn(n, seq_length)
    rec = np.random.rand(n, seq_length)
    chemo = np.random.rand(n, seq_length)
    radio = np.random.rand(n, seq_length)
    consts = dev.sim_consts(simulation_params['window_size'], simulation_params['lag'])
    params_dev = dev.to_device(dev.pack_params(simulation_params))
    aa = None if assigned_actions is None else dev.to_device(np.asarray(assigned_actions, dtype=np.float64))
    out, _ = dev.sim_factual(params_dev, dev.to_device(noise), dev.to_device(rec), dev.to_device(chemo),
                             dev.to_device(radio), int(seq_length), consts, assigned_actions=aa)
    torch.cuda.current_stream().synchronize()
    outputs = {k: out[k].cpu().numpy() for k in
               ('cancer_volume', 'chemo_dosage', 'radio_dosage', 'chemo_application', 'radio_application',
                'chemo_probabilities', 'radio_probabilities')}
    outputs['sequence_lengths'] = out['sequence_lengths'].cpu().numpy()
    outputs['death_flags'] = out['death_flags'].cpu().numpy()
    outputs['recovery_flags'] = out['recovery_flags'].cpu().numpy()
    outputs['patient_types'] = simulation_params['patient_types']
    assert not np.any(np.isnan(outputs['cancer_volume'])), 'Cancer volume contains NaN'
    return outputs


def simulate_counterfactual_1_step(simulation_params, seq_length, lazy=False):
    """All one-step-ahead counterfactuals of the test patients.  Reference: :378-563.
    lazy=True (used by this package's own dataset classes): the three dense (R, T) arrays are expanded and copied to
    the host on first access only; the dictionary carries the device-resident compact cohort."""
    from . import counterfactual as cf
    n = simulation_params['initial_stages'].shape[0]
    draws = _draw_per_patient(n, seq_length, 0)
    return cf.one_step_dense(simulation_params, int(seq_length), draws, lazy=lazy)


def simulate_counterfactuals_treatment_seq(simulation_params, seq_length, projection_horizon,
                                           cf_seq_mode='sliding_treatment', lazy=False):
    """Multi-step counterfactual treatment sequences of the test patients.  Reference: :566-773.  lazy: see above."""
    from . import counterfactual as cf
    if cf_seq_mode != 'sliding_treatment':
        # 'random_trajectories' consumes the RNG data-dependently inside the time loop (:704-705);
        # it is not configured anywhere in the reference (config/dataset/cancer_sim.yaml:17).
        raise NotImplementedError(f"cf_seq_mode={cf_seq_mode!r}: only 'sliding_treatment' is implemented")
    n = simulation_params['initial_stages'].shape[0]
    draws = _draw_per_patient(n, seq_length, int(projection_horizon))
    return cf.treatment_seq_dense(simulation_params, int(seq_length), int(projection_horizon), draws, lazy=lazy)


def get_scaling_params(sim):
    """mean / std (ddof=0) over the active entries, as two pandas Series.  Reference: :776-796.

    Boolean-mask indexing concatenates the active prefixes in row order, i.e. the same element order
    as the reference's python lists, so np.mean / np.std return bit-identical values."""
    seq = np.asarray(sim['sequence_lengths']).astype(np.int64)
    means, stds = {}, {}
    for k in ('cancer_volume', 'chemo_dosage', 'radio_dosage'):
        a = np.asarray(sim[k])
        active = a[np.arange(a.shape[1])[None, :] < seq[:, None]]
        means[k] = np.mean(active)
        stds[k] = np.std(active)
    means['patient_types'] = np.mean(sim['patient_types'])
    stds['patient_types'] = np.std(sim['patient_types'])
    return pd.Series(means), pd.Series(stds)
