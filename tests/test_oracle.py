"""CPU tests: the oracle against the reference's golden vectors (no GPU, no /root/reference needed).

Fixtures in tests/golden/ were produced by oracle/make_golden.py from the UNMODIFIED reference."""
import hashlib

import numpy as np
import pytest

from oracle import rng_export as rx
from oracle import sim_oracle as so
from oracle import sindy_np as sp
from oracle.ref_loader import reference_available

import helpers as h


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_np_mean_matches_numpy_bitwise():
    rng = np.random.RandomState(0)
    for _ in range(3000):
        n = rng.randint(1, 40)
        a = rng.rand(n) * 10 ** rng.uniform(-3, 3, size=n)
        assert np.array(list(a)).mean() == so.np_mean(a)


@pytest.fixture(scope="module")
def small():
    return h.load_npz('ref_sim_small.npz')


@pytest.fixture(scope="module")
def small_inputs():
    return h.collection_inputs(7, 2.0, 192, 24, 24)


def test_rng_replay_reproduces_reference_params(small, small_inputs):
    for name in ('train', 'val', 'one', 'seq'):
        p = small_inputs[name][0]
        for k in h.PARAM_KEYS:
            assert np.array_equal(p[k], small[f'{name}/params/{k}']), (name, k)
        assert np.array_equal(p['initial_stages'], small[f'{name}/params/initial_stages'])


def test_c_oracle_matches_reference_small(small, small_inputs):
    o = h.oracle_collection(small_inputs)
    exact = {'train': ['cancer_volume', 'chemo_dosage', 'radio_dosage', 'chemo_application', 'radio_application',
                       'sequence_lengths', 'death_flags', 'recovery_flags'],
             'one': ['chemo_application', 'radio_application', 'sequence_lengths', 'patient_types'],
             'seq': ['chemo_application', 'radio_application', 'sequence_lengths', 'patient_types',
                     'patient_ids_all_trajectories', 'patient_current_t']}
    exact['val'] = exact['train']
    for name, keys in exact.items():
        for k in keys:
            assert np.array_equal(o[name][k], small[f'{name}/out/{k}']), (name, k)
    # floats that go through numpy's SIMD exp/log differ from libm by <= a few ulp
    for name, k in (('train', 'chemo_probabilities'), ('train', 'radio_probabilities'), ('one', 'cancer_volume'),
                    ('seq', 'cancer_volume')):
        assert o[name][k].shape == small[f'{name}/out/{k}'].shape
        np.testing.assert_allclose(o[name][k], small[f'{name}/out/{k}'], rtol=1e-13, atol=1e-15)


def test_scaling_params_match_reference(small, small_inputs):
    o = so.sim_factual(small_inputs['train'][0], 60, small_inputs['train'][1])
    means, stds = so.scaling_params(o)
    order = ['cancer_volume', 'chemo_dosage', 'radio_dosage', 'patient_types']
    assert [means[k] for k in order] == list(small['train/scaling_means'])
    assert [stds[k] for k in order] == list(small['train/scaling_stds'])


def test_c_oracle_gamma10_and_assigned_actions():
    g = h.load_npz('ref_sim_gamma10.npz')
    np.random.seed(100)
    p = rx.generate_params(128, 10.0, 10.0, 15, 0)
    d = rx.draw_factual(128, 60)
    for k in h.PARAM_KEYS:
        assert np.array_equal(p[k], g[f'params/{k}'])
    o = so.sim_factual(p, 60, d)
    for k in ('cancer_volume', 'chemo_application', 'radio_application', 'sequence_lengths', 'death_flags',
              'recovery_flags', 'chemo_dosage'):
        assert np.array_equal(o[k], g[f'out/{k}']), k
    d2 = rx.draw_factual(128, 60)
    o2 = so.sim_factual(p, 60, d2, assigned_actions=g['assigned_actions'])
    for k in ('cancer_volume', 'chemo_application', 'radio_application', 'sequence_lengths', 'chemo_probabilities'):
        assert np.array_equal(o2[k], g[f'out_assigned/{k}']), k


@pytest.fixture(scope="module")
def seed1():
    inputs = h.collection_inputs(1, 2.0, 1000, 100, 100)
    return inputs, h.oracle_collection(inputs)


def test_seed1_digests(seed1):
    """Log configuration (seed 1, gamma 2, 1000/100/100): bit-exact arrays vs sha256 of the reference."""
    inputs, o = seed1
    dig = h.load_json('ref_digests_seed1.json')
    for name in ('train', 'val', 'one', 'seq'):
        for k in h.PARAM_KEYS:
            assert _digest(inputs[name][0][k]) == dig[name]['params_sha256'][k], (name, k)
        assert o[name]['cancer_volume'].shape[0] == dig[name]['rows']
        for k in ('chemo_application', 'radio_application', 'sequence_lengths'):
            assert _digest(o[name][k]) == dig[name]['out_sha256'][k], (name, k)
    assert dig['one']['rows'] == 22748 and dig['seq']['rows'] == 56239
    for k in ('cancer_volume', 'chemo_dosage', 'radio_dosage', 'death_flags', 'recovery_flags'):
        assert _digest(o['train'][k]) == dig['train']['out_sha256'][k], k
    assert _digest(o['one']['cancer_volume']) == dig['one']['out_sha256']['cancer_volume']
    assert abs(o['seq']['cancer_volume'].sum() - dig['seq']['cancer_volume_sum']) < 1e-6


def test_population_fit_and_metrics_reproduce_reference_log(seed1):
    """Known-answer test: results/2_main_table/final_with_insite.txt:6 of the reference."""
    _, o = seed1
    log = h.load_json('ref_log_seed1.json')['sindy']
    means, stds = so.scaling_params(o['train'])
    dtr, sc = sp.process_data(o['train'], means, stds)
    coefs, sup, stats = sp.fit_population(dtr, sc)
    np.testing.assert_allclose(coefs, np.array(log['coefs']), rtol=1e-10)
    assert sup.all()
    assert [s[1] for s in stats] == [42092, 20567, 20850, 10471]
    d1, _ = sp.process_data(o['one'], means, stds)
    orig, all_, last = sp.masked_rmse(sp.predictions_population(d1, sc, coefs), d1, sc)
    np.testing.assert_allclose([all_, orig, last], [log['encoder_test_rmse_all'], log['encoder_test_rmse_orig'],
                                                    log['encoder_test_rmse_last']], rtol=1e-11)
    d2, _ = sp.process_data(o['seq'], means, stds)
    d2s = sp.process_sequential_test(d2, sc, 5)
    ps = sp.slice_autoregressive(sp.predictions_population(d2, sc, coefs), d2['sequence_lengths'], 5)
    np.testing.assert_allclose(sp.n_step_rmses(ps, d2s, sc), log['decoder_test_rmse_2_to_6_step'], rtol=1e-11)
    assert sp.equation_string(coefs).startswith('Treatment 0: x_dot = +-0.0560145608')


def test_joint_model_reproduces_reference_ablation_log():
    """Known-answer test for the joint ("one ODE", 11-term) model: results/ablation/one_ode/...txt:10 of the reference
    (multilabel treatments; the run's cached dataset is the seed-10 collection)."""
    log = h.load_json('ref_log_joint_seed10.json')
    o = h.oracle_collection(h.collection_inputs(log['seed'], 2.0, 1000, 100, 100))
    means, stds = so.scaling_params(o['train'])
    dtr, sc = sp.process_data(o['train'], means, stds, treatment_mode='multilabel')
    coefs, sup, rows = sp.fit_population_joint(dtr, sc)
    np.testing.assert_allclose(coefs[0], log['sindy']['coefs'], rtol=1e-11)
    assert sup.all() and rows == log['sindy']['rows']
    assert sp.equation_string_joint(coefs)[:33] == log['sindy']['global_equation_string'][:33]   # same format, ~1e-14 digits
    d1, _ = sp.process_data(o['one'], means, stds, treatment_mode='multilabel')
    orig, all_, last = sp.masked_rmse(sp.predictions_population_joint(d1, sc, coefs), d1, sc)
    np.testing.assert_allclose([all_, orig, last], [log['sindy']['encoder_test_rmse_all'], log['sindy']['encoder_test_rmse_orig'],
                                                    log['sindy']['encoder_test_rmse_last']], rtol=1e-11)
    d2, _ = sp.process_data(o['seq'], means, stds, treatment_mode='multilabel')
    d2s = sp.process_sequential_test(d2, sc, 5)
    ps = sp.slice_autoregressive(sp.predictions_population_joint(d2, sc, coefs), d2['sequence_lengths'], 5)
    np.testing.assert_allclose(sp.n_step_rmses(ps, d2s, sc), log['sindy']['decoder_test_rmse_2_to_6_step'], rtol=1e-11)
    # the 11-term expression restricted to a treatment is the 4-term ODE the rollout kernels integrate
    c44 = sp.joint_to_per_treatment(sp.effective_coefs(coefs))
    codes = (np.squeeze(d1['current_treatments'])[..., 0] + 2 * np.squeeze(d1['current_treatments'])[..., 1]).astype(np.int64)
    prev = np.squeeze(d1['prev_outputs'] * sc['output_stds'] + sc['output_means'], -1)
    static = d1['static_features'] * sc['inputs_stds'][1:2] + sc['input_means'][1:2]
    un = sp.rollout_unscaled(prev[:, 0], codes, static[:, 0], c44)
    ref = sp.predictions_population_joint(d1, sc, coefs)[..., 0] * sc['output_stds'] + sc['output_means']
    np.testing.assert_allclose(un, ref, rtol=1e-9, atol=1e-9)


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted (GPU box)")
def test_oracle_against_live_reference():
    """In the build container the unmodified reference is importable: compare directly."""
    import warnings
    from oracle.ref_loader import load_reference_sim
    warnings.filterwarnings('ignore')
    m = load_reference_sim()
    np.random.seed(3)
    p = m.generate_params(40, 2.0, 2.0, 15, 0)
    st = np.random.get_state()
    ref = m.simulate_counterfactuals_treatment_seq(p, 60, 5)
    np.random.set_state(st)
    d = rx.draw_cf(40, 60, 5)
    o = so.sim_cf_treatment_seq(p, 60, 5, d)
    assert o['cancer_volume'].shape == ref['cancer_volume'].shape
    for k in ('chemo_application', 'radio_application', 'sequence_lengths', 'patient_ids_all_trajectories',
              'patient_current_t'):
        assert np.array_equal(o[k], ref[k]), k
    np.testing.assert_allclose(o['cancer_volume'], ref['cancer_volume'], rtol=1e-13, atol=1e-15)


def test_philox_block_function_known_answers():
    """Random123 kat_vectors for philox4x32-10 pin the generator restatement (oracle/philox_np.py)."""
    from oracle import philox_np as ph
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = ph.philox4x32_10(*[np.uint64(c) for c in ctr], *key)
        assert tuple(int(g) for g in got) == want
    d = ph.draw_factual(2000, 60, 3)
    assert abs(d['noise'].std() / 0.01 - 1) < 0.01 and abs(d['noise'].mean()) < 1e-4
    for k in ('recovery', 'chemo', 'radio'):
        assert 0.0 <= d[k].min() and d[k].max() < 1.0 and abs(d[k].mean() - 0.5) < 0.005
    assert abs(np.corrcoef(d['noise'][:, 0::2].ravel(), d['noise'][:, 1::2].ravel())[0, 1]) < 0.01


@pytest.mark.parametrize("kind", ['one', 'seq'])
def test_windowed_oracle_equals_the_verbatim_restatement(kind):
    """oracle_sim_cf_*_windowed (a patient's treatment window read from a given output row instead of the running
    output array) is what the depth-4 GPU tests use to check single patients deep inside 10^5..10^6-patient cohorts;
    fed the rows of the verbatim restatement it must reproduce that restatement bit for bit."""
    from oracle import sim_oracle as so
    n = 500
    params, draws = h.random_cohort(n, seed=9, extra=5 if kind == 'seq' else 0)
    run = (lambda p, d, **kw: so.sim_cf_treatment_seq(p, 60, 5, d, **kw)) if kind == 'seq' else \
          (lambda p, d, **kw: so.sim_cf_one_step(p, 60, d, **kw))
    full = run(params, draws)
    idx = np.arange(1, n)
    sub = {k: (v[idx] if isinstance(v, np.ndarray) else v) for k, v in params.items()}
    part = run(sub, {k: v[idx] for k, v in draws.items()}, window_rows=full['cancer_volume'][idx])
    r0 = full['cancer_volume'].shape[0] - part['cancer_volume'].shape[0]
    assert r0 > 0
    for k in ('cancer_volume', 'chemo_application', 'radio_application', 'sequence_lengths'):
        assert np.array_equal(full[k][r0:], part[k], equal_nan=True), k


def test_odeint_restatement_known_answers_of_the_reference():
    """The reference's own in-file tests of odeint (pkpd/utils.py:759-780 dense grid, :807-828 two-point grid):
    dy/dt = 1 from y0 = 0 gives y = t, MSE < 1e-16, with hmax = HMAX = STANDARD_DT / STEPS_FOR_DT."""
    HMAX = sp.STANDARD_DT / sp.STEPS_FOR_DT
    t = np.arange(0, 10.0, 10.0 / 60)
    ones = lambda y, h: np.ones_like(y)
    y = sp.odeint_euler(ones, np.array(0.0), t, hmax=HMAX)
    assert y.shape == t.shape and np.mean((y - t) ** 2) < 1e-16
    t2 = np.array([t[0], t[-1]])
    y2 = sp.odeint_euler(ones, np.array(0.0), t2, hmax=HMAX)
    assert np.mean((y2 - t2) ** 2) < 1e-16
    # the rollout restatement on the same grids: coefficient row [1, 0, 0, 0] is dy/dt = 1
    c = np.zeros((4, 4)); c[:, 0] = 1.0
    codes = np.zeros((1, len(t) - 1), dtype=np.int64)
    r = sp.rollout_unscaled(np.zeros(1), codes, np.ones(1), c, dts=np.diff(t))
    assert np.mean((r[0] - t[1:]) ** 2) < 1e-16
    # uniform interval lengths reproduce the uniform-dt code path bit for bit
    rng = np.random.RandomState(0)
    cc = rng.randn(4, 4) * 0.1
    cd = rng.randint(0, 4, size=(5, 20))
    x0 = rng.rand(5) * 100
    u = rng.randint(1, 4, size=5).astype(float)
    a = sp.rollout_unscaled(x0, cd, u, cc)
    b = sp.rollout_unscaled(x0, cd, u, cc, dts=np.full(20, sp.STANDARD_DT))
    assert np.array_equal(a, b)


def test_lsq_initial_mask_restatement_is_pinned_by_the_reference_log(seed1):
    """The restatement of the reference's dormant per-patient optimiser LSQIntialMask (pkpd/utils.py:244-327) with a
    dense warm start and the unbias refit IS pysindy's STLSQ, so on the seed-1 training data it has to reproduce the 16
    logged population coefficients (final_with_insite.txt:6); a sparse warm start only restricts the support."""
    _, o = seed1
    log = np.array(h.load_json('ref_log_seed1.json')['sindy']['coefs'])
    means, stds = so.scaling_params(o['train'])
    dtr, sc = sp.process_data(o['train'], means, stds)
    buckets = sp.de_format_snippets(dtr, sc)
    for a in range(4):
        th, xd = sp.design_matrices(buckets[a])
        c, ind = sp.lsq_initial_mask(th, xd, np.ones(4), 1e-3, 0.5, unbias=True)
        np.testing.assert_allclose(c, log[a], rtol=1e-10)
        assert ind.all()
        guess = log[a].copy(); guess[3] = 0.0
        c2, ind2 = sp.lsq_initial_mask(th, xd, guess, 1e-3, 0.5, unbias=False)
        assert not ind2[3] and c2[3] == 0.0 and ind2[:3].all()
        from sklearn.linear_model import ridge_regression
        np.testing.assert_allclose(c2[:3], ridge_regression(th[:, :3], xd, 0.5, tol=1e-6), rtol=1e-12)



def test_population_fit_second_known_answer_seed10():
    """A second pin of the restated population path: final_with_insite.txt:2094 (SINDy on the collection seeded with 10):
    coefficients parsed from the logged equation string and the 8 RMSEs."""
    import ast
    import re
    ref = ast.literal_eval(h.load_json('ref_logline_seed1.json')['sindy_seed10'])
    o = h.oracle_collection(h.collection_inputs(10, 2.0, 1000, 100, 100))
    means, stds = so.scaling_params(o['train'])
    dtr, sc = sp.process_data(o['train'], means, stds)
    coefs, sup, _ = sp.fit_population(dtr, sc)
    logged = [float(c) for c in re.findall(r'\+(-?[0-9.]+(?:e-?[0-9]+)?)\*', ref['global_equation_string'])]
    assert len(logged) == 16
    np.testing.assert_allclose(coefs.reshape(-1), logged, rtol=1e-9)
    d1, _ = sp.process_data(o['one'], means, stds)
    orig, all_, last = sp.masked_rmse(sp.predictions_population(d1, sc, coefs), d1, sc)
    np.testing.assert_allclose([all_, orig, last], [ref['encoder_test_rmse_all'], ref['encoder_test_rmse_orig'],
                                                    ref['encoder_test_rmse_last']], rtol=1e-10)
    d2, _ = sp.process_data(o['seq'], means, stds)
    d2s = sp.process_sequential_test(d2, sc, 5)
    ps = sp.slice_autoregressive(sp.predictions_population(d2, sc, coefs), d2['sequence_lengths'], 5)
    np.testing.assert_allclose(sp.n_step_rmses(ps, d2s, sc), [ref[f'decoder_test_rmse_{k}-step'] for k in range(2, 7)],
                               rtol=1e-10)


def test_savgol_window2_restatement_equals_scipy():
    """pysindy SmoothedFiniteDifference(smoother_kws={'window_length': 2, 'polyorder': 1}) (sindy.py:196-198) calls
    scipy.signal.savgol_filter(x, axis=0): scipy is installed here, so the restatement is pinned on the real thing."""
    from scipy.signal import savgol_filter
    rng = np.random.default_rng(3)
    for L in (2, 3, 4, 5, 9, 31, 60):
        x = rng.uniform(0.1, 1200.0, size=(L, 1))
        np.testing.assert_allclose(sp.savgol_w2_p1(x[:, 0]), savgol_filter(x, window_length=2, polyorder=1, axis=0)[:, 0],
                                   rtol=1e-13)


def test_degree4_library_order_and_names_equal_sklearn():
    """pysindy's PolynomialLibrary subclasses sklearn's PolynomialFeatures (sindy.py:185-186 passes degree=4,
    interaction_only=False): the column order and the feature names of the restatement are pinned on the real class."""
    from sklearn.preprocessing import PolynomialFeatures
    rng = np.random.default_rng(11)
    X = np.stack([rng.uniform(0.1, 30.0, 200), rng.integers(1, 4, 200).astype(np.float64)], axis=1)
    pf = PolynomialFeatures(degree=4, interaction_only=False, include_bias=True).fit(X)
    np.testing.assert_allclose(sp.library_poly4(X[:, 0], X[:, 1]), pf.transform(X), rtol=1e-14)
    assert tuple(map(tuple, pf.powers_)) == sp.POLY4_EXPONENTS
    assert tuple(pf.get_feature_names_out(['x0', 'u0'])) == sp.POLY4_NAMES
    # degree 2, interaction only: the 4-term and 11-term libraries of the default and the joint model
    pf2 = PolynomialFeatures(degree=2, interaction_only=True, include_bias=True).fit(X)
    np.testing.assert_allclose(sp.library_p4(X[:, 0], X[:, 1]), pf2.transform(X), rtol=1e-15)
    X4 = np.concatenate([X, rng.integers(0, 2, (200, 2)).astype(np.float64)], axis=1)[:, [0, 2, 3, 1]]
    pf11 = PolynomialFeatures(degree=2, interaction_only=True, include_bias=True).fit(X4)
    np.testing.assert_allclose(sp.library_p11(X4[:, 0], X4[:, 1:]), pf11.transform(X4), rtol=1e-15)
    assert tuple(n.replace(' ', '*') for n in pf11.get_feature_names_out(['x0', 'u0', 'u1', 'u2'])) == sp.JOINT_NAMES


def test_degree4_fit_is_insensitive_to_the_unbias_cutoff(seed1):
    """The degree-4 library is rank 12 of 15 on cancer_sim data; the fit must not depend on which of the two historical
    cut-offs of the un-bias step (scipy's machine epsilon / today's scikit-learn 1e-6) is used, nor on the row order."""
    _, o = seed1
    means, stds = so.scaling_params(o['train'])
    dtr, sc = sp.process_data(o['train'], means, stds)
    buckets = sp.de_format_snippets(dtr, sc)
    th, xd = sp.design_matrices_poly4(buckets[3])
    assert np.linalg.matrix_rank(th) == 12
    c_eps, s_eps = sp.stlsq_fit(th, xd, 1e-3, 0.5, scipy_lstsq=True)
    c_skl, s_skl = sp.stlsq_fit(th, xd, 1e-3, 0.5, scipy_lstsq=False)
    assert np.array_equal(s_eps, s_skl) and 5 <= s_eps.sum() < 15
    np.testing.assert_allclose(c_eps, c_skl, rtol=1e-9, atol=1e-14)
    perm = np.random.default_rng(0).permutation(th.shape[0])
    c_perm, s_perm = sp.stlsq_fit(th[perm], xd[perm], 1e-3, 0.5, scipy_lstsq=True)
    assert np.array_equal(s_perm, s_eps)
    np.testing.assert_allclose(c_perm, c_eps, rtol=1e-9, atol=1e-14)
