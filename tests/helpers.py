"""Shared helpers of the test-suite: seeded cohorts through the oracle (tests only)."""
import json
import os
import warnings

import numpy as np

from oracle import rng_export as rx
from oracle import sim_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PARAM_KEYS = ['patient_types', 'initial_volumes', 'alpha', 'rho', 'beta', 'beta_c', 'K',
              'chemo_sigmoid_intercepts', 'radio_sigmoid_intercepts', 'chemo_sigmoid_betas', 'radio_sigmoid_betas']


def load_npz(name):
    return np.load(os.path.join(GOLD, name))


def load_json(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def collection_inputs(seed, gamma, n_train, n_val, n_test, T=60, H=5):
    """Parameters and random draws of the four subsets in the reference's RNG order
    (dataset.py:589-601): returns dict name -> (params, draws)."""
    warnings.filterwarnings('ignore')
    np.random.seed(seed)
    out = {}
    p = rx.generate_params(n_train, gamma, gamma, 15, 0); out['train'] = (p, rx.draw_factual(n_train, T))
    p = rx.generate_params(n_val, gamma, gamma, 15, 0); out['val'] = (p, rx.draw_factual(n_val, T))
    p = rx.generate_params(n_test, gamma, gamma, 15, 0); out['one'] = (p, rx.draw_cf(n_test, T))
    p = rx.generate_params(n_test, gamma, gamma, 15, 0); out['seq'] = (p, rx.draw_cf(n_test, T, H))
    return out


def oracle_collection(inputs, T=60, H=5):
    return {'train': so.sim_factual(inputs['train'][0], T, inputs['train'][1]),
            'val': so.sim_factual(inputs['val'][0], T, inputs['val'][1]),
            'one': so.sim_cf_one_step(inputs['one'][0], T, inputs['one'][1]),
            'seq': so.sim_cf_treatment_seq(inputs['seq'][0], T, H, inputs['seq'][1])}


def random_cohort(n, seed, gamma=2.0, T=60, extra=0, per_patient=False):
    """Cheap seeded cohort for large-N tests: reference parameter generator + vectorised draws
    (NOT the reference's draw order; used where only self-consistency / oracle parity matters)."""
    warnings.filterwarnings('ignore')
    np.random.seed(seed)
    p = rx.generate_params(n, gamma, gamma, 15, 0)
    rng = np.random.RandomState(seed + 1)
    draws = dict(noise=0.01 * rng.randn(n, T + extra), recovery=rng.rand(n, T), chemo=rng.rand(n, T),
                 radio=rng.rand(n, T))
    return p, draws


def max_rel(a, b, floor=1e-300):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor))) if a.size else 0.0
