"""TEST INFRASTRUCTURE ONLY.  Replays the reference's global-numpy-RNG draw order.

The reference consumes ``np.random`` (legacy MT19937 global state) in this order
(cancer_simulation.py):

  get_standard_params :96-215   choice(stages,N,p) :112 -> truncnorm.rvs per sorted stage :135 ->
                                multivariate_normal(size=N) repeated until N accepted :163-173 ->
                                choice([1,2,3],N) :177 -> truncnorm.rvs(size=N) :190-193 ->
                                shuffle(idx) :209
  simulate_factual              randn(N,T), rand(N,T) x3 (noise, recovery, chemo, radio) :275-279
  simulate_counterfactual_1_step        per patient: randn(T), rand(T), rand(T), rand(T) :440-453
  simulate_counterfactuals_treatment_seq per patient: randn(T+H), rand(T), rand(T), rand(T) :640-653

The functions below draw from the *current global state*, so ``np.random.seed(s)`` followed by
the same sequence of calls the reference makes yields the reference's arrays bit for bit
(checked in tests/test_oracle.py against the reference itself).
"""
import numpy as np
from scipy.stats import truncnorm

# cancer_simulation.py:47-59
_SIZE_DIST = {'I': (1.72, 4.70, 0.3, 5.0), 'II': (1.96, 1.63, 0.3, 13.0), 'IIIA': (1.91, 9.40, 0.3, 13.0),
              'IIIB': (2.76, 6.87, 0.3, 13.0), 'IV': (3.86, 8.82, 0.3, 13.0)}
_STAGE_OBS = {'I': 1432, 'II': 128, 'IIIA': 1306, 'IIIB': 7248, 'IV': 12840}


def calc_volume(diameter):
    return 4 / 3 * np.pi * (diameter / 2) ** 3          # :34-35


def calc_diameter(volume):
    return ((volume / (4 / 3 * np.pi)) ** (1 / 3)) * 2  # :38-39


TUMOUR_DEATH_THRESHOLD = calc_volume(13)                # :44


def generate_params(num_patients, chemo_coeff, radio_coeff, window_size, lag):
    """Restatement of generate_params/get_standard_params (:66-215), same RNG consumption."""
    n = num_patients
    total = sum(_STAGE_OBS.values())
    stages = sorted(_SIZE_DIST)
    initial_stages = np.random.choice(stages, n, p=[_STAGE_OBS[s] / total for s in stages])
    diam, stage_names = [], []
    for s in stages:
        cnt = int(np.sum((initial_stages == s) * 1))
        mu, sigma, lo, hi = _SIZE_DIST[s]
        lo = (np.log(lo) - mu) / sigma
        hi = (np.log(hi) - mu) / sigma
        rv = truncnorm.rvs(lo, hi, size=cnt)
        diam += list(np.exp((rv * sigma) + mu))
        stage_names += [s] * cnt
    K = calc_volume(30)
    rho_p, alpha_p, beta_c_p = (7 * 10 ** -5, 7.23 * 10 ** -3), (0.0398, 0.168), (0.028, 0.0007)
    c = 0.87 * alpha_p[1] * rho_p[1]
    cov = np.array([[alpha_p[1] ** 2, c], [c, rho_p[1] ** 2]])
    mean = np.array([alpha_p[0], rho_p[0]])
    acc = []
    while len(acc) < n:
        holder = np.random.multivariate_normal(mean, cov, size=n)
        for i in range(holder.shape[0]):
            if holder[i, 0] > 0.0 and holder[i, 1] > 0.0:
                acc.append(holder[i, :])
    patient_types = np.random.choice([1, 2, 3], n)
    chemo_adj = np.array([0.0 if i < 3 else 0.1 for i in patient_types])
    radio_adj = np.array([0.0 if i > 1 else 0.1 for i in patient_types])
    acc = np.array(acc)[:n, :]
    alpha = acc[:, 0] + alpha_p[0] * radio_adj
    rho = acc[:, 1]
    beta = alpha / 10
    beta_c = beta_c_p[0] + beta_c_p[1] * truncnorm.rvs((0.0 - beta_c_p[0]) / beta_c_p[1],
                                                       (np.inf - beta_c_p[0]) / beta_c_p[1],
                                                       size=n) + beta_c_p[0] * chemo_adj
    holder = {'patient_types': patient_types, 'initial_stages': np.array(stage_names),
              'initial_volumes': calc_volume(np.array(diam)), 'alpha': alpha, 'rho': rho, 'beta': beta,
              'beta_c': beta_c, 'K': np.array([K for _ in range(n)])}
    idx = [i for i in range(n)]
    np.random.shuffle(idx)
    out = {k: v[idx] for k, v in holder.items()}
    d_max = calc_diameter(TUMOUR_DEATH_THRESHOLD)
    out['chemo_sigmoid_intercepts'] = np.array([d_max / 2.0 for _ in patient_types])
    out['radio_sigmoid_intercepts'] = np.array([d_max / 2.0 for _ in patient_types])
    out['chemo_sigmoid_betas'] = np.array([chemo_coeff / d_max for _ in patient_types])
    out['radio_sigmoid_betas'] = np.array([radio_coeff / d_max for _ in patient_types])
    out['window_size'] = window_size
    out['lag'] = lag
    return out


def draw_factual(n, T):
    """:275-279 -> dict(noise (already x0.01), recovery, chemo, radio), each (n,T)."""
    noise = 0.01 * np.random.randn(n, T)
    rec = np.random.rand(n, T)
    chemo = np.random.rand(n, T)
    radio = np.random.rand(n, T)
    return dict(noise=noise, recovery=rec, chemo=chemo, radio=radio)


def draw_cf(n, T, extra=0):
    """:440-453 (extra=0) / :640-653 (extra=H): per-patient draws, stacked to (n,T+extra)/(n,T)."""
    noise = np.empty((n, T + extra)); rec = np.empty((n, T)); chemo = np.empty((n, T)); radio = np.empty((n, T))
    for i in range(n):
        noise[i] = 0.01 * np.random.randn(T + extra)
        rec[i] = np.random.rand(T)
        chemo[i] = np.random.rand(T)
        radio[i] = np.random.rand(T)
    return dict(noise=noise, recovery=rec, chemo=chemo, radio=radio)
