"""GPU tests of the compact-cohort evaluation path (K8 cf_eval_one_step, K9 cf_eval_treatment_seq, the per-(patient, t)
individualisation kernels): BASELINE config C3 without dense rows.

Parity statements
  * population SINDy: the 8 RMSEs of the reference's committed run log (final_with_insite.txt:6) come out of the compact
    path at rel 1e-9, from the logged coefficients and the seed-1 test cohorts in their compact form;
  * compact == dense: the error sums equal b200i_masked_se over the expanded reference rows rolled out by K6 (rel 1e-10),
    for population coefficients and for per-(patient, t) coefficients, on cohorts with deaths, recoveries, NaN-dropped
    options and a patient that stops at t = 0;
  * one fit per (patient, t): b200i_insite_bfgs_prefix / b200i_stlsq_prefix return, for every dense row, the
    coefficients b200i_insite_bfgs / b200i_stlsq_batched compute on that row (bit-identical inputs -> rel 1e-12);
  * INSITE through the compact path: the 8 RMSEs of the reference's INSITE log line (:2362) at 1e-9 (jax's BFGS
    restated with its failure semantics; that run predates the zoom-failure fallback, see tests/test_gpu_insite.py)."""
import numpy as np
import pytest

import helpers as h

pytestmark = pytest.mark.gpu
T, H = 60, 5


@pytest.fixture(scope="module")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from b200_insite import device
    device.require_cuda()
    return device


def _cohorts(dev, params_draws_one, params_draws_seq):
    from b200_insite import counterfactual as cf
    out = []
    for kind, (params, draws) in (('one', params_draws_one), ('seq', params_draws_seq)):
        pd_ = dev.to_device(dev.pack_params(params))
        dd = [dev.to_device(draws[k]) for k in ('noise', 'recovery', 'chemo', 'radio')]
        static = dev.to_device(np.asarray(params['patient_types'], dtype=np.float64))
        cohort = cf.sim_cf_one_step(pd_, *dd, T) if kind == 'one' else cf.sim_cf_treatment_seq(pd_, *dd, T, H)
        out.append((cohort, static))
    return out


@pytest.fixture(scope="module")
def seed1(dev):
    inputs = h.collection_inputs(1, 2.0, 1000, 100, 100)
    return _cohorts(dev, inputs['one'], inputs['seq'])


def test_population_rmses_of_the_reference_log_from_compact_cohorts(dev, seed1):
    from b200_insite import compact_eval as ce
    log = h.load_json('ref_log_seed1.json')['sindy']
    coefs = dev.to_device(np.array(log['coefs']))
    (one, s1), (seq, s2) = seed1
    assert one.total_rows == 22748 and seq.total_rows == 56239
    res = ce.evaluate_model(one, s1, seq, s2, coefs, insite=False)
    for k in ('encoder_test_rmse_all', 'encoder_test_rmse_orig', 'encoder_test_rmse_last'):
        np.testing.assert_allclose(res[k], log[k], rtol=1e-9, err_msg=k)
    got = [res[f'decoder_test_rmse_{k}-step'] for k in range(2, 7)]
    np.testing.assert_allclose(got, log['decoder_test_rmse_2_to_6_step'], rtol=1e-9)


def _dense_rows(dev, cohort, static):
    """Reference rows of a compact cohort as the inputs of the dense kernels: x (R,W) = prev_outputs, targets (R,W) =
    outputs, codes (R,W) = chemo + 2*radio of the first W application columns, sequence lengths, static feature."""
    import torch
    from b200_insite import counterfactual as cf
    dense = cf.expand(cohort, static)
    vol = dense['cancer_volume']
    W = vol.shape[1] - 1
    x = vol[:, :W].contiguous()
    y = vol[:, 1:].contiguous()
    codes = dev.treatment_codes(dense['chemo_application'], dense['radio_application'], W)
    seq = dense['sequence_lengths'].to(torch.int32)
    return x, y, codes, seq, dense['patient_types'].contiguous(), dense


def _dense_sums(dev, cohort, static, coefs_rows, drop_below):
    """masked_se sums of the dense path: K6 over the expanded rows, then the metric masks of the reference."""
    import torch
    x, y, codes, seq, st, dense = _dense_rows(dev, cohort, static)
    pred = dev.ode_rollout(x[:, 0].contiguous(), st, codes, coefs_rows, drop_below=drop_below)
    if cohort.kind == 'one_step':
        return dev.masked_se(pred, y, seq).cpu().numpy()
    R, W = pred.shape
    lo = torch.clamp(torch.clamp(seq.to(torch.int64) - H, min=1), max=W - H)          # sindy.py:729-733
    idx = lo[:, None] + torch.arange(H, device='cuda')[None, :]
    p5, y5 = torch.gather(pred, 1, idx), torch.gather(y, 1, idx)
    se = ((p5 - y5) ** 2).sum(0).cpu().numpy()
    return np.concatenate([se, np.full(H, float(R))])


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


@pytest.mark.parametrize("n,seed", [(1, 3), (7, 4), (150, 11), (333, 12)])
def test_compact_sums_equal_the_dense_path(dev, n, seed):
    import torch
    from b200_insite import compact_eval as ce
    coefs = dev.to_device(np.array(h.load_json('ref_log_seed1.json')['sindy']['coefs']))
    p1, d1 = h.random_cohort(n, seed=seed)
    p2, d2 = h.random_cohort(n, seed=seed + 100, extra=H)
    rng = np.random.RandomState(seed)
    for cohort, static in _cohorts(dev, (p1, d1), (p2, d2)):
        got = ce.evaluate(cohort, static, coefs, 1e-3).cpu().numpy()
        ref = _dense_sums(dev, cohort, static, coefs, 1e-3)
        assert got.shape == ref.shape
        W = T - 1
        if cohort.kind == 'one_step':
            assert np.array_equal(got[W:2 * W], ref[W:2 * W]), "active rows per column"
            assert got[3 * W + 1] == ref[3 * W + 1] == cohort.total_rows
        else:
            assert np.array_equal(got[H:], ref[H:]) and got[H] == cohort.total_rows
        assert _rel(got, ref) < 1e-10, (cohort.kind, n)
        # per-(patient, t) coefficients: random matrices near the population's, gathered per dense row for K6
        per = coefs.reshape(1, 1, 4, 4) * (1.0 + 0.05 * dev.to_device(rng.randn(n, W, 4, 4)))
        per = per.contiguous()
        x, y, codes, seq, st, dense = _dense_rows(dev, cohort, static)
        off = cohort.row_offsets.cpu().numpy()
        rows = np.arange(cohort.total_rows)
        owner = np.searchsorted(off, rows, side='right') - 1
        sl = dense['sequence_lengths'].cpu().numpy().astype(np.int64)
        t_of = sl - 1 if cohort.kind == 'one_step' else sl - H - 1
        per_rows = per[torch.from_numpy(owner).cuda(), torch.from_numpy(t_of).cuda()].contiguous()
        got = ce.evaluate(cohort, static, per, -1.0).cpu().numpy()
        ref = _dense_sums(dev, cohort, static, per_rows, -1.0)
        assert _rel(got, ref) < 1e-10, (cohort.kind, n, 'per step')


def test_empty_cohort_and_bad_arguments(dev):
    import torch
    from b200_insite import _native
    f64 = dict(dtype=torch.float64, device='cuda')
    F = torch.zeros((0, T), **f64)
    codes = torch.zeros((0, T), dtype=torch.uint8, device='cuda')
    ns = torch.zeros((0,), dtype=torch.int32, device='cuda')
    st = torch.zeros((0,), **f64)
    coefs = torch.zeros((4, 4), **f64)
    s = dev.cf_eval_one_step(F, codes, torch.zeros((0, T - 1, 4), **f64), ns, st, coefs).cpu().numpy()
    assert s.shape == (3 * (T - 1) + 2,) and not s.any()
    s = dev.cf_eval_treatment_seq(F, codes, torch.zeros((0, T - 1, 2 * H, H), **f64),
                                  torch.zeros((0, T - 1), dtype=torch.int16, device='cuda'), ns, st, coefs).cpu().numpy()
    assert s.shape == (2 * H,) and not s.any()
    lib = _native.load()
    rc = lib.b200i_cf_eval_one_step(4, 300, 0.1, 5, None, None, None, None, None, None, 0, 1e-3, dev._ptr(coefs), None)
    assert rc == -2 and b'outside' in lib.b200i_last_error()
    rc = lib.b200i_cf_eval_treatment_seq(4, 60, 9, 0.1, 5, None, None, None, None, None, None, None, 0, 1e-3,
                                         dev._ptr(coefs), None)
    assert rc == -2
    rc = lib.b200i_cf_eval_one_step(4, 60, 0.1, 5, None, None, None, None, None, None, 0, 1e-3, dev._ptr(coefs), None)
    assert rc == -1 and b'NULL' in lib.b200i_last_error()


@pytest.mark.parametrize("estimator", ['bfgs_rollout', 'ridge_prior_stlsq'])
def test_one_fit_per_patient_step_equals_the_per_row_fits(dev, estimator):
    import torch
    from b200_insite import compact_eval as ce
    theta0 = dev.to_device(np.array(h.load_json('ref_log_seed1.json')['sindy']['coefs']))
    n = 40
    p1, d1 = h.random_cohort(n, seed=21)
    p2, d2 = h.random_cohort(n, seed=22, extra=H)
    for cohort, static in _cohorts(dev, (p1, d1), (p2, d2)):
        ph = 1 if cohort.kind == 'one_step' else H
        coefs, diag = ce.individualise(cohort, static, theta0, estimator=estimator, lam=10.0, ridge_prior_lam=1e4,
                                       zoom_failure_fallback=False)
        x, y, codes, seq, st, dense = _dense_rows(dev, cohort, static)
        W = x.shape[1]
        if estimator == 'bfgs_rollout':
            dense_coefs, dense_status, dense_fval = dev.insite_bfgs(x, codes, seq, ph, st, theta0, lam=10.0)
        else:
            fit_len = torch.clamp(seq - ph, 0, W - 1).to(torch.int32)
            dense_coefs = dev.stlsq_batched(x, codes, fit_len, st, theta0, lam=1e4)
        off = cohort.row_offsets.cpu().numpy()
        owner = np.searchsorted(off, np.arange(cohort.total_rows), side='right') - 1
        sl = dense['sequence_lengths'].cpu().numpy().astype(np.int64)
        t_of = sl - 1 if cohort.kind == 'one_step' else sl - H - 1
        got = coefs[torch.from_numpy(owner).cuda(), torch.from_numpy(t_of).cuda()].cpu().numpy()
        ref = dense_coefs.cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-14, err_msg=f"{estimator} {cohort.kind}")
        if estimator == 'bfgs_rollout':
            st_got = diag['status'][torch.from_numpy(owner).cuda(), torch.from_numpy(t_of).cuda()].cpu().numpy()
            assert np.array_equal(st_got, dense_status.cpu().numpy())
            # steps that were never executed keep theta0 and are marked skipped
            ns = cohort.n_steps.cpu().numpy()
            full = diag['status'].cpu().numpy()
            for i in range(n):
                assert (full[i, ns[i]:] == -2).all()
        # and the metrics through per-step coefficients equal the dense rollout with the per-row coefficients
        got_s = ce.evaluate(cohort, static, coefs, -1.0).cpu().numpy()
        ref_s = _dense_sums(dev, cohort, static, dense_coefs, -1.0)
        assert _rel(got_s, ref_s) < 1e-10


def test_insite_rmses_of_the_reference_log_from_compact_cohorts(dev, seed1):
    """INSITE (per-row BFGS, lam 10) on the seed-1 test cohorts: 11 700 fits instead of 78 987."""
    from b200_insite import compact_eval as ce
    log = h.load_json('ref_log_seed1.json')
    theta0 = dev.to_device(np.array(log['sindy']['coefs']))
    (one, s1), (seq, s2) = seed1
    res = ce.evaluate_model(one, s1, seq, s2, theta0, insite=True, estimator='bfgs_rollout', lam=10.0,
                            zoom_failure_fallback=False)
    ins = log['insite']
    for k in ('encoder_test_rmse_all', 'encoder_test_rmse_orig', 'encoder_test_rmse_last'):
        np.testing.assert_allclose(res[k], ins[k], rtol=1e-9, err_msg=k)
    got = [res[f'decoder_test_rmse_{k}-step'] for k in range(2, 7)]
    np.testing.assert_allclose(got, ins['decoder_test_rmse_2_to_6_step'], rtol=1e-9)
