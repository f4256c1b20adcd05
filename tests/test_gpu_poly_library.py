"""model.ablation_more_complex_basis_functions (sindy.py:185-186): the degree-4 library of the four per-treatment
population models.  pysindy is not installed here; the oracle restates the library and the STLSQ loop and calls the real
sklearn ridge_regression / scipy.linalg.lstsq, which is what pysindy's optimiser runs on (parity otherwise unpinned: the
reference holds no logged run with this flag).  Tolerances: support bit-exact, coefficients 1e-7 relative (the library is
rank 12 of 15 with singular values over 13 orders of magnitude; the oracle itself moves by 1e-12 under a row permutation)."""
import numpy as np
import pytest

import helpers as h

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from b200_insite import device
    device.require_cuda()
    return device


@pytest.fixture(scope="module")
def collection():
    from b200_insite.dataset import SyntheticCancerDatasetCollection
    col = SyntheticCancerDatasetCollection(2.0, 2.0, {'train': 1000, 'val': 100, 'test': 100}, seed=1)
    col.process_data_multi()
    return col


@pytest.fixture(scope="module")
def fitted(dev, collection):
    from b200_insite.config import default_config
    from b200_insite.sindy import SINDY
    model = SINDY(default_config(insite=False, ablation_more_complex_basis_functions=True), collection)
    model.fit(collection.train_f)
    return model


@pytest.fixture(scope="module")
def oracle_fit(collection):
    from oracle import sindy_np as sp
    tr = collection.train_f
    return sp.fit_population_poly4(tr.data, tr.scaling_params)


def test_r_factors_carry_the_design_matrices(fitted, oracle_fit):
    """R^T R == [Theta | xdot]^T [Theta | xdot] entry by entry relative to the column norms, counts exact."""
    _, _, mats = oracle_fit
    rf = fitted.population_stats_
    for a in range(4):
        th, xd = mats[a]
        M = np.concatenate([th, xd[:, None]], axis=1)
        R = rf[a * 256:(a + 1) * 256].reshape(16, 16)
        assert np.all(np.tril(R, -1) == 0.0)
        assert rf[4 * 256 + a] == M.shape[0]
        G, Gr = M.T @ M, R.T @ R
        d = np.sqrt(np.diag(G))
        np.testing.assert_allclose(Gr / np.outer(d, d), G / np.outer(d, d), atol=1e-11)
        # and the singular values the minimum-norm step depends on
        s_ref = np.linalg.svd(th, compute_uv=False)
        s_dev = np.linalg.svd(R[:15, :15], compute_uv=False)
        # (LAPACK's own absolute error on the 42092 x 15 matrix is ~ eps * s_max = 2e-3; the factor's columns are exact
        # relative to their own norms, which is what the scaled comparison above checks)
        np.testing.assert_allclose(s_dev[:12], s_ref[:12], rtol=1e-9, atol=4 * 2.3e-16 * s_ref[0])
        assert s_dev[12] < 1e-9 * s_dev[0]                      # rank 12 of 15: three patient types


def test_population_fit_equals_oracle(fitted, oracle_fit):
    coefs, sup, _ = oracle_fit
    assert fitted.joint_coefs.shape == (4, 15)
    assert np.array_equal(fitted.support_, sup)
    assert 5 <= sup.sum(1).min() and sup.sum(1).max() < 15      # x0^2 .. x0^4 terms fall below the threshold
    np.testing.assert_allclose(fitted.joint_coefs, coefs, rtol=1e-7, atol=1e-13)
    assert fitted.feature_library_names[4] == 'x0 u0' and fitted.feature_library_names[-1] == 'u0^4'
    s = fitted.global_equation_string
    assert s.startswith('Treatment 0: x_dot = +') and 'static_feature' not in s and '*u0^2' in s.replace('**', '^')


def test_rollout_and_metrics_equal_oracle(dev, fitted, collection):
    """The eight RMSEs of a run with the flag against the oracle's own pipeline on the oracle's own data."""
    from oracle import sim_oracle as so, sindy_np as sp
    from b200_insite.sindy import run_experiment
    from b200_insite.config import default_config
    res, model = run_experiment(default_config(insite=False, ablation_more_complex_basis_functions=True), collection)
    np.testing.assert_array_equal(model.joint_coefs, fitted.joint_coefs)
    o = h.oracle_collection(h.collection_inputs(1, 2.0, 1000, 100, 100))
    means, stds = so.scaling_params(o['train'])
    d1, sc = sp.process_data(o['one'], means, stds)
    pred = sp.predictions_population_poly4(d1, sc, fitted.joint_coefs)
    np.testing.assert_allclose(model.get_predictions(collection.test_cf_one_step), pred, rtol=1e-9, atol=1e-9)
    orig, all_, last = sp.masked_rmse(pred, d1, sc)
    np.testing.assert_allclose([res['encoder_test_rmse_all'], res['encoder_test_rmse_orig'], res['encoder_test_rmse_last']],
                               [all_, orig, last], rtol=1e-8)
    d2, _ = sp.process_data(o['seq'], means, stds)
    d2s = sp.process_sequential_test(d2, sc, 5)
    ps = sp.slice_autoregressive(sp.predictions_population_poly4(d2, sc, fitted.joint_coefs), d2['sequence_lengths'], 5)
    np.testing.assert_allclose([res[f'decoder_test_rmse_{k}-step'] for k in range(2, 7)], sp.n_step_rmses(ps, d2s, sc),
                               rtol=1e-8)


def test_degree4_is_refused_where_it_is_not_built(collection):
    from b200_insite.config import default_config
    from b200_insite.sindy import SINDY
    with pytest.raises(NotImplementedError):
        SINDY(default_config(insite=True, ablation_more_complex_basis_functions=True), collection)
    with pytest.raises(NotImplementedError):
        SINDY(default_config(insite=False, ablation_more_complex_basis_functions=True, joint_model=True,
                             treatment_mode='multilabel'), collection)


def test_edge_cases_empty_and_ragged_cohorts(dev):
    """No patients, one patient, a treatment that never occurs, one-step trajectories: counts, zero factors and zero
    coefficients where there is nothing to fit; the R factors still reproduce the explicit design matrix."""
    import torch
    from oracle import sindy_np as sp
    T = 12
    rng = np.random.default_rng(5)

    def run(vol, chemo, radio, seq, static):
        rf = dev.poly_tsqr(dev.to_device(vol), dev.to_device(chemo), dev.to_device(radio), dev.to_device(seq),
                           dev.to_device(static))
        coefs, sup = dev.poly_stlsq(rf)
        torch.cuda.synchronize()
        return rf.cpu().numpy(), coefs.cpu().numpy(), sup.cpu().numpy()

    z = np.zeros((0, T))
    rf, coefs, sup = run(z, z, z, np.zeros(0), np.zeros(0))
    assert not rf.any() and not coefs.any() and not sup.any()
    # ragged cohort: radiotherapy never given, sequence lengths 1 .. T-1
    n = 37
    vol = rng.uniform(0.5, 30.0, size=(n, T))
    chemo = (rng.uniform(size=(n, T)) < 0.4).astype(np.float64)
    radio = np.zeros((n, T))
    seq = rng.integers(1, T, size=n).astype(np.float64)
    seq[0], seq[1] = 1, T - 1
    static = rng.integers(1, 4, size=n).astype(np.float64)
    rf, coefs, sup = run(vol, chemo, radio, seq, static)
    # explicit design matrices from the same snippet rule (per-treatment runs, end point with the backward difference)
    rows = {0: [], 1: []}
    for p in range(n):
        L, a = int(seq[p]), 0
        for i in range(1, L + 1):
            if i == L or chemo[p, i] != chemo[p, i - 1]:
                x = vol[p, a:i + 1]
                th = sp.library_poly4(x, np.full(len(x), static[p]))
                xd = sp.finite_difference_order1(x, sp.STANDARD_DT)
                rows[int(chemo[p, a])].append(np.concatenate([th, xd[:, None]], axis=1))
                a = i
    for a in (0, 1):
        M = np.concatenate(rows[a], axis=0)
        R = rf[a * 256:(a + 1) * 256].reshape(16, 16)
        assert rf[4 * 256 + a] == M.shape[0]
        G = M.T @ M
        d = np.sqrt(np.diag(G))
        np.testing.assert_allclose((R.T @ R) / np.outer(d, d), G / np.outer(d, d), atol=1e-11)
    for a in (2, 3):     # radiotherapy never occurs
        assert rf[4 * 256 + a] == 0 and not rf[a * 256:(a + 1) * 256].any()
        assert not coefs[a].any() and not sup[a].any()
    assert np.isfinite(coefs).all()
    # one patient, one step
    rf, coefs, sup = run(vol[:1], chemo[:1], radio[:1], np.ones(1), static[:1])
    assert rf[4 * 256:].sum() == 2 and np.isfinite(coefs).all()
