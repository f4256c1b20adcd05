"""GPU tests of irregular sampling (BASELINE config C4): the `_dts` entry points of K4, K5b, K6, K7.

  * known answers of the reference: its in-file odeint tests (pkpd/utils.py:759-828) -- dy/dt = 1 => y = t on the dense
    60-point grid and on the two-point grid [t0, t_end], MSE < 1e-16 -- through b200i_ode_rollout_dts;
  * uniform interval lengths reproduce the uniform entry points bit for bit (K4, K5b, K6 FP64/FP32, K7);
  * irregular grids (shared and per-row) against the numpy restatement (oracle/sindy_np.py): K6 1e-12, FP32 1e-4,
    K5b 1e-7 with bit-exact support, K4 normal equations 1e-11 / sample counts exact, K7 reported objective = restated
    objective at the returned point (1e-9) and optimum <= scipy's;
  * FP32 storage of the volumes in K5b: 1e-4 from FP64 (north star's FP32 tolerance)."""
import numpy as np
import pytest

import helpers as h

pytestmark = pytest.mark.gpu
DT = 10.0 / 60


@pytest.fixture(scope="module")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from b200_insite import device
    device.require_cuda()
    return device


@pytest.fixture(scope="module")
def cohort():
    """Factual cohort through the oracle: x = volumes, codes, sequence lengths, static feature."""
    from oracle import sim_oracle as so
    params, draws = h.random_cohort(700, seed=5)
    sim = so.sim_factual(params, 60, draws)
    vol = sim['cancer_volume']
    codes = (sim['chemo_application'] + 2 * sim['radio_application']).astype(np.uint8)
    return dict(vol=vol, chemo=sim['chemo_application'], radio=sim['radio_application'], codes=codes,
                seq=sim['sequence_lengths'].astype(np.int64), u=np.asarray(params['patient_types'], dtype=np.float64))


def _grids(R, W, seed):
    rng = np.random.RandomState(seed)
    shared = DT * rng.uniform(0.3, 2.5, size=W)
    per_row = DT * rng.uniform(0.3, 2.5, size=(R, W))
    return shared, per_row


def test_reference_odeint_known_answers(dev):
    import torch
    t = np.arange(0, 10.0, DT)                      # the reference's dense grid (60 points)
    c = np.zeros((4, 4)); c[:, 0] = 1.0             # dy/dt = 1
    for grid in (t, np.array([t[0], t[-1]])):       # dense and two-point ("sparse") grids
        dts = dev.intervals_from_times(dev.to_device(grid))
        W = len(grid) - 1
        for fp32 in (False, True):
            y = dev.ode_rollout(dev.to_device(np.zeros(3)), dev.to_device(np.ones(3)),
                                torch.zeros((3, W), dtype=torch.uint8, device='cuda'), dev.to_device(c), drop_below=-1.0,
                                dts=dts, fp32=fp32).cpu().numpy()
            mse = np.mean((y - grid[None, 1:]) ** 2)
            assert mse < (1e-10 if fp32 else 1e-16), (len(grid), fp32, mse)


def test_uniform_interval_lengths_reproduce_the_uniform_entry_points(dev, cohort):
    import torch
    c = cohort
    R, T = c['vol'].shape
    W = T - 1
    theta0 = np.array(h.load_json('ref_log_seed1.json')['sindy']['coefs'])
    x = dev.to_device(np.ascontiguousarray(c['vol'][:, :W]))
    cd = dev.to_device(np.ascontiguousarray(c['codes'][:, :W]), dtype=torch.uint8)
    u = dev.to_device(c['u'])
    seq = dev.to_device(c['seq'], dtype=torch.int32)
    fit_len = dev.to_device(np.clip(c['seq'] - 1, 0, W - 1), dtype=torch.int32)
    prior = dev.to_device(theta0)
    for dts in (torch.full((W,), DT, dtype=torch.float64, device='cuda'),
                torch.full((R, W), DT, dtype=torch.float64, device='cuda')):
        per = dev.stlsq_batched(x, cd, fit_len, u, prior, 1e4)
        assert torch.equal(per, dev.stlsq_batched(x, cd, fit_len, u, prior, 1e4, dts=dts))
        for fp32 in (False, True):
            a = dev.ode_rollout(x[:, 0].contiguous(), u, cd, per, drop_below=-1.0, fp32=fp32)
            b = dev.ode_rollout(x[:, 0].contiguous(), u, cd, per, drop_below=-1.0, fp32=fp32, dts=dts)
            assert torch.equal(a, b)
        n7 = 64
        a = dev.insite_bfgs(x[:n7].contiguous(), cd[:n7].contiguous(), seq[:n7].contiguous(), 1, u[:n7].contiguous(), prior, 10.0)
        b = dev.insite_bfgs(x[:n7].contiguous(), cd[:n7].contiguous(), seq[:n7].contiguous(), 1, u[:n7].contiguous(), prior, 10.0,
                            dts=dts[:n7].contiguous() if dts.dim() == 2 else dts)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # K4: Gram part of the population statistics
    vol, ch, ra = (dev.to_device(c[k]) for k in ('vol', 'chemo', 'radio'))
    sl = dev.to_device(c['seq'].astype(np.float64))
    ref = dev.theta_gram(vol, ch, ra, sl, u).cpu().numpy()[:60]
    got = dev.theta_gram_dts(vol, ch, ra, sl, u, torch.full((R, T - 1), DT, dtype=torch.float64, device='cuda')).cpu().numpy()[:60]
    np.testing.assert_allclose(got, ref, rtol=1e-12)       # same samples, different reduction shape


@pytest.mark.parametrize("per_row", [False, True])
def test_irregular_rollout_and_fits_match_the_numpy_restatement(dev, cohort, per_row):
    import torch
    from oracle import sindy_np as sp
    c = cohort
    R, T = c['vol'].shape
    W = T - 1
    theta0 = np.array(h.load_json('ref_log_seed1.json')['sindy']['coefs'])
    theta0[3, 3] = 5e-4                                   # one off-support term
    shared, rows = _grids(R, W, 3)
    dts_np = rows if per_row else shared
    dts = dev.to_device(dts_np)
    x_np = np.ascontiguousarray(c['vol'][:, :W]); cd_np = np.ascontiguousarray(c['codes'][:, :W])
    x = dev.to_device(x_np); cd = dev.to_device(cd_np, dtype=torch.uint8); u = dev.to_device(c['u'])
    # K6, population and per-row coefficients
    rng = np.random.RandomState(1)
    per_np = theta0[None] * (1 + 0.05 * rng.randn(R, 4, 4))
    for coefs_np, drop in ((theta0, 1e-3), (per_np, -1.0)):
        ce = np.where(np.abs(coefs_np) > drop, coefs_np, 0.0)
        ref = sp.rollout_unscaled(x_np[:, 0], cd_np.astype(np.int64), c['u'], ce, dts=dts_np)
        got = dev.ode_rollout(x[:, 0].contiguous(), u, cd, dev.to_device(coefs_np), drop_below=drop, dts=dts)
        assert h.max_rel(got.cpu().numpy(), ref, floor=1e-6) < 1e-12
        got32 = dev.ode_rollout(x[:, 0].contiguous(), u, cd, dev.to_device(coefs_np), drop_below=drop, dts=dts, fp32=True)
        scale = np.abs(ref).max()
        assert np.max(np.abs(got32.cpu().numpy() - ref)) / scale < 1e-4      # north star: 1e-4 in FP32
    # K5b
    fit_len = np.clip(c['seq'] - 1, 0, W - 1).astype(np.int32)
    for lam, thr in ((1e4, 1e-3), (1.0, 0.05)):
        got = dev.stlsq_batched(x, cd, dev.to_device(fit_len, dtype=torch.int32), u, dev.to_device(theta0), lam,
                                threshold=thr, dts=dts).cpu().numpy()
        for r in range(0, R, 7):
            ref = sp.ridge_prior_row(x_np[r], cd_np[r], c['u'][r], fit_len[r], theta0, lam, threshold=thr,
                                     dts=dts_np[r] if per_row else dts_np)
            assert np.array_equal(got[r] != 0, ref != 0), r
            np.testing.assert_allclose(got[r], ref, rtol=1e-7, atol=1e-10)
    # K4
    vol, ch, ra = (dev.to_device(c[k]) for k in ('vol', 'chemo', 'radio'))
    stats = dev.theta_gram_dts(vol, ch, ra, dev.to_device(c['seq'].astype(np.float64)), u, dts).cpu().numpy()
    un = dev.unpack_stats(stats)
    n4 = 120
    stats4 = dev.theta_gram_dts(vol[:n4].contiguous(), ch[:n4].contiguous(), ra[:n4].contiguous(),
                                dev.to_device(c['seq'][:n4].astype(np.float64)), u[:n4].contiguous(),
                                dts[:n4].contiguous() if per_row else dts).cpu().numpy()
    un4 = dev.unpack_stats(stats4)
    G, b, cnt = sp.normal_equations_irregular(c['vol'][:n4], c['chemo'][:n4], c['radio'][:n4], c['seq'][:n4], c['u'][:n4],
                                              dts_np[:n4] if per_row else dts_np)
    assert np.array_equal(un4['count'], cnt)
    np.testing.assert_allclose(un4['G'], G, rtol=1e-11)
    np.testing.assert_allclose(un4['b'], b, rtol=1e-9, atol=1e-9 * np.abs(b).max())
    assert un['count'].sum() > cnt.sum()
    # K7: objective parity at the returned point and optimum <= scipy's BFGS on the same objective
    n7 = 24
    theta0 = np.array(h.load_json('ref_log_seed1.json')['sindy']['coefs'])
    coefs, status, fval = dev.insite_bfgs(x[:n7].contiguous(), cd[:n7].contiguous(),
                                          dev.to_device(c['seq'][:n7], dtype=torch.int32), 1, u[:n7].contiguous(),
                                          dev.to_device(theta0), 10.0,
                                          dts=dts[:n7].contiguous() if per_row else dts, line_search='robust')
    coefs, fval = coefs.cpu().numpy(), fval.cpu().numpy()
    from scipy.optimize import minimize
    for r in range(n7):
        n_fit = min(int(c['seq'][r]) - 1, W - 1)
        if n_fit <= 0:
            continue
        d = dts_np[r] if per_row else dts_np
        start = sp.insite_objective(theta0.reshape(-1), x_np[r], cd_np[r], c['u'][r], n_fit, theta0.reshape(-1), 10.0, 1.0,
                                    with_grad=False, dts=d)
        norm = 2.5 * start
        fe = sp.insite_objective(coefs[r].reshape(-1), x_np[r], cd_np[r], c['u'][r], n_fit, theta0.reshape(-1), 10.0, norm,
                                 with_grad=False, dts=d)
        np.testing.assert_allclose(fval[r, 1], fe, rtol=1e-9)
        res = minimize(lambda th: sp.insite_objective(th, x_np[r], cd_np[r], c['u'][r], n_fit, theta0.reshape(-1), 10.0, norm,
                                                      dts=d), theta0.reshape(-1), jac=True, method='BFGS',
                       options={'gtol': 1e-12, 'maxiter': 3200})
        assert fe <= res.fun * (1 + 1e-6) + 1e-12, (r, fe, res.fun)


def test_fp32_storage_of_the_volumes_in_the_batched_fit(dev, cohort):
    import torch
    c = cohort
    R, T = c['vol'].shape
    W = T - 1
    theta0 = np.array(h.load_json('ref_log_seed1.json')['sindy']['coefs'])
    x = dev.to_device(np.ascontiguousarray(c['vol'][:, :W]))
    cd = dev.to_device(np.ascontiguousarray(c['codes'][:, :W]), dtype=torch.uint8)
    u = dev.to_device(c['u'])
    fit_len = dev.to_device(np.clip(c['seq'] - 1, 0, W - 1), dtype=torch.int32)
    prior = dev.to_device(theta0)
    ref = dev.stlsq_batched(x, cd, fit_len, u, prior, 1e4)
    got = dev.stlsq_batched(x.to(torch.float32).contiguous(), cd, fit_len, u, prior, 1e4)
    scale = ref.abs().amax(dim=(1, 2), keepdim=True)
    assert float(((got - ref).abs() / scale).max().item()) < 1e-4
    assert torch.equal(got != 0, ref != 0)
