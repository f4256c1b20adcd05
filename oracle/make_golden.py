"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/* by running the UNMODIFIED reference
simulator (imported from /root/reference through oracle/ref_loader.py) in the build container.

    python -m oracle.make_golden

Outputs (committed):
  tests/golden/ref_sim_small.npz      dense reference outputs of all four generators for a small
                                      collection (seed 7, gamma 2, 192/24/24/24 patients) + the
                                      parameter dicts the reference drew.
  tests/golden/ref_sim_gamma10.npz    simulate_factual, seed 100, gamma 10, 128 patients (the
                                      reference's own __main__ setting, cancer_simulation.py:844-852,
                                      at reduced N) incl. an `assigned_actions` fixed-policy run.
  tests/golden/ref_digests_seed1.json sha256 digests + row counts of the reference outputs for the
                                      log's configuration (seed 1, gamma 2, 1000/100/100), and the
                                      get_scaling_params values.
  tests/golden/ref_log_seed1.json     known-answer numbers copied from the reference's committed
                                      run log results/2_main_table/final_with_insite.txt:6 / :2362.
"""
import hashlib
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference_sim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
_PARAM_KEYS = ['patient_types', 'initial_volumes', 'alpha', 'rho', 'beta', 'beta_c', 'K',
               'chemo_sigmoid_intercepts', 'radio_sigmoid_intercepts', 'chemo_sigmoid_betas',
               'radio_sigmoid_betas']


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _collection(m, seed, gamma, n_train, n_val, n_test, T=60, H=5):
    np.random.seed(seed)
    out = {}
    p = m.generate_params(n_train, gamma, gamma, 15, 0); out['train'] = (p, m.simulate_factual(p, T))
    p = m.generate_params(n_val, gamma, gamma, 15, 0); out['val'] = (p, m.simulate_factual(p, T))
    p = m.generate_params(n_test, gamma, gamma, 15, 0); out['one'] = (p, m.simulate_counterfactual_1_step(p, T))
    p = m.generate_params(n_test, gamma, gamma, 15, 0)
    out['seq'] = (p, m.simulate_counterfactuals_treatment_seq(p, T, H))
    return out


def main():
    warnings.filterwarnings('ignore')
    m = load_reference_sim()
    os.makedirs(GOLD, exist_ok=True)

    # 1. small dense collection ---------------------------------------------------------------
    col = _collection(m, 7, 2.0, 192, 24, 24)
    blob = {}
    for name, (p, d) in col.items():
        for k in _PARAM_KEYS:
            blob[f'{name}/params/{k}'] = p[k]
        blob[f'{name}/params/initial_stages'] = p['initial_stages'].astype('U4')
        for k, v in d.items():
            blob[f'{name}/out/{k}'] = v
    means, stds = m.get_scaling_params(col['train'][1])
    blob['train/scaling_means'] = means.values
    blob['train/scaling_stds'] = stds.values
    np.savez_compressed(os.path.join(GOLD, 'ref_sim_small.npz'), **blob)

    # 2. gamma = 10 factual + fixed-policy run ------------------------------------------------
    np.random.seed(100)
    p = m.generate_params(128, 10.0, 10.0, 15, 0)
    d = m.simulate_factual(p, 60)
    st = np.random.get_state()
    aa = np.random.RandomState(5).rand(128, 60, 2)
    d2 = m.simulate_factual(p, 60, assigned_actions=aa)
    blob = {f'params/{k}': p[k] for k in _PARAM_KEYS}
    blob.update({f'out/{k}': v for k, v in d.items()})
    blob.update({f'out_assigned/{k}': v for k, v in d2.items()})
    blob['assigned_actions'] = aa
    np.savez_compressed(os.path.join(GOLD, 'ref_sim_gamma10.npz'), **blob)

    # 3. digests of the log's configuration ---------------------------------------------------
    col = _collection(m, 1, 2.0, 1000, 100, 100)
    dig = {'config': {'seed': 1, 'gamma': 2.0, 'n_train': 1000, 'n_val': 100, 'n_test': 100, 'T': 60, 'H': 5}}
    for name, (p, d) in col.items():
        dig[name] = {'rows': int(d['cancer_volume'].shape[0]),
                     'params_sha256': {k: _digest(p[k]) for k in _PARAM_KEYS},
                     'out_sha256': {k: _digest(v) for k, v in d.items()},
                     'cancer_volume_sum': float(d['cancer_volume'].sum()),
                     'sequence_lengths_sum': float(d['sequence_lengths'].sum())}
    means, stds = m.get_scaling_params(col['train'][1])
    dig['train']['scaling_means'] = {k: float(v) for k, v in means.items()}
    dig['train']['scaling_stds'] = {k: float(v) for k, v in stds.items()}
    with open(os.path.join(GOLD, 'ref_digests_seed1.json'), 'w') as f:
        json.dump(dig, f, indent=1)

    # 4. known answers from the committed reference run log -------------------------------------
    log = {
        'source': 'results/2_main_table/final_with_insite.txt:6 (sindy) and :2362 (insite), seed 1, gamma 2, '
                  '1000/100/100 patients, multiclass, sliding_treatment',
        'sindy': {
            'coefs': [[-0.05601456082026624, -0.11598756834077, -0.07958279124512227, 0.07326347275734027],
                      [-0.5517350343589641, -0.8761667536689084, -0.053397817822270766, -0.035996455669168224],
                      [-3.649800303098579, -0.8626472911638889, 1.1157911611997717, -0.6373790072514276],
                      [-1.6336216074116419, -3.49858670473956, -3.584018882175004, 0.06151172618047967]],
            'encoder_test_rmse_all': 2.156088252028878,
            'encoder_test_rmse_orig': 1.7342646495065481,
            'encoder_test_rmse_last': 1.7533225309220122,
            'decoder_test_rmse_2_to_6_step': [1.327252086844152, 1.3053306069334676, 1.297414430898069,
                                              1.2942172534259109, 1.289684544561717]},
        'insite': {
            'encoder_test_rmse_all': 1.0837636799825472,
            'encoder_test_rmse_orig': 0.828552914710425,
            'encoder_test_rmse_last': 1.0039120469218676,
            'decoder_test_rmse_2_to_6_step': [0.7955476123367343, 0.7807888962696772, 0.7790102860782369,
                                              0.7843825525412129, 0.7881425166336781]}}
    with open(os.path.join(GOLD, 'ref_log_seed1.json'), 'w') as f:
        json.dump(log, f, indent=1)
    for fn in sorted(os.listdir(GOLD)):
        print(fn, os.path.getsize(os.path.join(GOLD, fn)))


if __name__ == '__main__':
    main()
