"""Drop-in replacement for the reference module ``src.data.continuous.continuous`` (libs_m/ct/src/data/continuous/
continuous.py), the simulators behind the EQ_5_A..D datasets: the cancer simulator with

  * an ``equation`` argument that selects the patient population -- one patient type for EQ_5_A / EQ_5_B, three for
    EQ_5_C / EQ_5_D, and a per-patient chemo sensitivity beta_c only for EQ_5_D (:200-216);
  * observation noise ``0.01 * np.random.normal`` added to the volumes of EQ_5_B / C / D after the simulation, over the
    WHOLE pre-allocated buffer of the counterfactual simulators (:373-374, :568-569, :784-785), so the global RNG is
    consumed exactly as the reference consumes it;
  * a continuous dose channel: ``chemo_dosage`` rows in the counterfactual dictionaries (:571, :788).

    generate_params(num_patients, chemo_coeff, radio_coeff, window_size, lag, equation)                 <- :68-95, :98-224
    simulate_factual(simulation_params, seq_length, equation, assigned_actions=None)                    <- :227-388
    simulate_counterfactual_1_step(simulation_params, seq_length, equation)                             <- :391-582
    simulate_counterfactuals_treatment_seq(simulation_params, seq_length, projection_horizon, equation,
                                           cf_seq_mode='sliding_treatment')                             <- :585-800
    get_scaling_params(sim)                                                                             <- :803-824

The time stepping itself is the cancer simulator's, so it runs in the same CUDA kernels (K1 / K2 / K3, csrc/sim_*.cu):
the cross-row treatment window of the counterfactual simulators reads the buffers before the noise is added (:471 of
the cancer simulator = :486 here), exactly as the kernels model it.
"""
import enum

import numpy as np
from scipy.stats import truncnorm

from . import cancer_simulation as cs
from .cancer_simulation import (TUMOUR_CELL_DENSITY, TUMOUR_DEATH_THRESHOLD, calc_diameter, calc_volume,  # noqa: F401
                                cancer_stage_observations, get_scaling_params, tumour_size_distributions)

OBSERVATION_NOISE = 0.01


class Equation(enum.IntEnum):
    """src.data.pkpd.pkpd_simulation.Equation (:51-60); only the name is read by the simulators."""
    EQ_4_A = 1
    EQ_4_B = 2
    EQ_4_C = 3
    EQ_4_D = 4
    EQ_5_A = 5
    EQ_5_B = 6
    EQ_5_C = 7
    EQ_5_D = 8
    EQ_4_M = 9


def _name(equation):
    name = equation.name if hasattr(equation, 'name') else str(equation)
    if name not in ('EQ_5_A', 'EQ_5_B', 'EQ_5_C', 'EQ_5_D'):
        raise ValueError(f"equation {name!r}: the continuous simulators serve EQ_5_A .. EQ_5_D")
    return name


def _noisy(name):
    return name.split('_')[-1] in ('B', 'C', 'D')


def get_standard_params(num_patients, equation):
    """Static per-patient parameters; RNG draw order of the reference (:121, :143, :181, :197, :210, :230)."""
    name = _name(equation)
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
    while n_kept < n:
        draw = np.random.multivariate_normal(mean, cov, size=n)
        ok = draw[(draw[:, 0] > 0.0) & (draw[:, 1] > 0.0)]
        kept.append(ok)
        n_kept += ok.shape[0]
    alpha_rho = np.concatenate(kept, axis=0)[:n] if kept else np.zeros((0, 2))

    possible_types = [1] if name in ('EQ_5_A', 'EQ_5_B') else [1, 2, 3]
    patient_types = np.random.choice(possible_types, n)
    chemo_adj = np.where(patient_types < 3, 0.0, 0.1)
    radio_adj = np.where(patient_types > 1, 0.0, 0.1)
    alpha = alpha_rho[:, 0] + alpha_mean * radio_adj
    rho = alpha_rho[:, 1]
    beta = alpha / 10
    beta_c_adj = beta_c_mean * chemo_adj
    if name == 'EQ_5_D':
        beta_c = beta_c_mean + beta_c_sd * truncnorm.rvs((0.0 - beta_c_mean) / beta_c_sd, (np.inf - beta_c_mean) / beta_c_sd,
                                                         size=n) + beta_c_adj
    else:
        beta_c = beta_c_mean + beta_c_adj
    holder = {'patient_types': patient_types, 'initial_stages': stage_labels, 'initial_volumes': calc_volume(diameters),
              'alpha': alpha, 'rho': rho, 'beta': beta, 'beta_c': beta_c, 'K': np.full(n, calc_volume(30))}
    order = list(range(n))
    np.random.shuffle(order)
    out = {k: v[order] for k, v in holder.items()}
    out['observation_noise'] = OBSERVATION_NOISE
    return out


def generate_params(num_patients, chemo_coeff, radio_coeff, window_size, lag, equation):
    params = get_standard_params(num_patients, equation)
    n = params['patient_types'].shape[0]
    d_max = calc_diameter(TUMOUR_DEATH_THRESHOLD)
    params['chemo_sigmoid_intercepts'] = np.full(n, d_max / 2.0)
    params['radio_sigmoid_intercepts'] = np.full(n, d_max / 2.0)
    params['chemo_sigmoid_betas'] = np.full(n, chemo_coeff / d_max)
    params['radio_sigmoid_betas'] = np.full(n, radio_coeff / d_max)
    params['window_size'] = window_size
    params['lag'] = lag
    return params


def simulate_factual(simulation_params, seq_length, equation, assigned_actions=None):
    """Factual trajectories (:227-388): the cancer simulator's kernel K1, then the observation noise."""
    name = _name(equation)
    out = cs.simulate_factual(simulation_params, seq_length, assigned_actions=assigned_actions)
    if _noisy(name):
        out['cancer_volume'] = out['cancer_volume'] + simulation_params['observation_noise'] * \
            np.random.normal(size=out['cancer_volume'].shape)
    keys = ('cancer_volume', 'chemo_dosage', 'radio_dosage', 'chemo_application', 'radio_application',
            'chemo_probabilities', 'radio_probabilities', 'sequence_lengths', 'death_flags', 'recovery_flags',
            'patient_types')
    out = {k: out[k] for k in keys}
    assert not np.any(np.isnan(out['cancer_volume'])), 'Cancer volume contains NaN'
    return out


def _chemo_dosage_rows(chemo_application, chemo_amt=5.0, decay=float(np.exp(-np.log(2) / 1))):
    """The dose channel of the exploded rows: every row carries the applications that were (factual part) or would be
    (its option) given, and the dosage obeys C[k] = C[k-1] * decay + chemo_amt * application[k] along it (:478-480,
    :526-527, :738-739); decay = exp(-ln 2) is exactly 0.5."""
    d = np.zeros_like(chemo_application)
    prev = np.zeros(chemo_application.shape[0])
    for k in range(chemo_application.shape[1]):
        prev = prev * decay + chemo_amt * chemo_application[:, k]
        d[:, k] = prev
    return d


def simulate_counterfactual_1_step(simulation_params, seq_length, equation):
    """One-step counterfactuals (:391-582).  The dose row of a counterfactual row ends at its own step t (:548); of a
    factual snapshot row it runs to t as well (the snapshot is taken after step t)."""
    name = _name(equation)
    n = simulation_params['initial_stages'].shape[0]
    base = cs.simulate_counterfactual_1_step(simulation_params, seq_length)
    seq = base['sequence_lengths'].astype(np.int64)
    dose = _chemo_dosage_rows(base['chemo_application'])
    dose[np.arange(dose.shape[1])[None, :] >= seq[:, None]] = 0.0      # rows are zero behind their last written step
    vol = base['cancer_volume']
    if _noisy(name):
        noise = np.random.normal(size=(n * int(seq_length) * 4, int(seq_length)))     # the whole buffer (:437, :568)
        vol = vol + simulation_params['observation_noise'] * noise[:vol.shape[0]]
    out = {'cancer_volume': vol, 'chemo_dosage': dose, 'chemo_application': base['chemo_application'],
           'radio_application': base['radio_application'], 'sequence_lengths': base['sequence_lengths'],
           'patient_types': base['patient_types']}
    assert not np.any(np.isnan(out['cancer_volume'])), 'Cancer volume contains NaN'
    return out


def simulate_counterfactuals_treatment_seq(simulation_params, seq_length, projection_horizon, equation,
                                           cf_seq_mode='sliding_treatment'):
    """Treatment-sequence counterfactuals (:585-800); the dose row covers t + 1 + projection_horizon steps (:771)."""
    name = _name(equation)
    n = simulation_params['initial_stages'].shape[0]
    H = int(projection_horizon)
    base = cs.simulate_counterfactuals_treatment_seq(simulation_params, seq_length, H, cf_seq_mode)
    seq = base['sequence_lengths'].astype(np.int64)
    dose = _chemo_dosage_rows(base['chemo_application'])
    dose[np.arange(dose.shape[1])[None, :] >= seq[:, None]] = 0.0
    vol = base['cancer_volume']
    if _noisy(name):
        noise = np.random.normal(size=(2 * H * n * int(seq_length), int(seq_length) + H))     # (:641, :784)
        vol = vol + simulation_params['observation_noise'] * noise[:vol.shape[0]]
    out = {'cancer_volume': vol, 'chemo_dosage': dose, 'chemo_application': base['chemo_application'],
           'radio_application': base['radio_application'], 'sequence_lengths': base['sequence_lengths'],
           'patient_types': base['patient_types'],
           'patient_ids_all_trajectories': base['patient_ids_all_trajectories'],
           'patient_current_t': base['patient_current_t']}
    assert not np.any(np.isnan(out['cancer_volume'])), 'Cancer volume contains NaN'
    return out
