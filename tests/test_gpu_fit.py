"""GPU parity tests of the SINDy fit-and-predict kernels (K4, K5, K6, metrics) against the oracle
and the reference's committed run log.  Support bit-exact; coefficients / RMSEs within 1e-6
relative (FP64 tolerance of BASELINE.json) -- asserted tighter where the conditioning allows."""
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
def seed1():
    from oracle import sim_oracle as so
    inputs = h.collection_inputs(1, 2.0, 1000, 100, 100)
    o = h.oracle_collection(inputs)
    means, stds = so.scaling_params(o['train'])
    return inputs, o, means, stds


def _oracle_gram(train):
    """Per-treatment normal equations from the oracle's explicit design matrices."""
    from oracle import sim_oracle as so, sindy_np as sp
    means, stds = so.scaling_params(train)
    dtr, sc = sp.process_data(train, means, stds)
    buckets = sp.de_format_snippets(dtr, sc)
    G, b, cnt = np.zeros((4, 4, 4)), np.zeros((4, 4)), np.zeros(4)
    for a in range(4):
        if buckets[a]:
            th, xd = sp.design_matrices(buckets[a])
            G[a], b[a], cnt[a] = th.T @ th, th.T @ xd, th.shape[0]
    return G, b, cnt


def _device_stats(dev, sim, static):
    import torch
    stats = dev.theta_gram(dev.to_device(sim['cancer_volume']), dev.to_device(sim['chemo_application']),
                           dev.to_device(sim['radio_application']), dev.to_device(sim['sequence_lengths']),
                           dev.to_device(np.asarray(static, dtype=np.float64)),
                           dev.to_device(sim['chemo_dosage']), dev.to_device(sim['radio_dosage']))
    torch.cuda.synchronize()
    return stats


def test_theta_gram_matches_oracle_design_matrices(dev, seed1):
    _, o, means, stds = seed1
    stats = _device_stats(dev, o['train'], o['train']['patient_types'])
    u = dev.unpack_stats(stats.cpu().numpy())
    G, b, cnt = _oracle_gram(o['train'])
    assert np.array_equal(u['count'], cnt)
    assert list(cnt) == [42092, 20567, 20850, 10471]
    np.testing.assert_allclose(u['G'], G, rtol=1e-11)
    np.testing.assert_allclose(u['b'], b, rtol=1e-9, atol=1e-6)
    sc = dev.moments_to_scaling(stats.cpu().numpy())
    for k in ('cancer_volume', 'chemo_dosage', 'radio_dosage'):
        np.testing.assert_allclose(sc[k], (means[k], stds[k]), rtol=1e-11)
    assert u['patients'] == 1000 and u['active'] == o['train']['sequence_lengths'].sum()


@pytest.mark.parametrize("pitch", [60, 64])
def test_lean_fit_side_outputs_reproduce_the_statistics(dev, pitch):
    """b200i_sim_factual_side + b200i_theta_gram_codes (treatment-code bytes and per-patient moment sums written by
    the simulator kernel) vs b200i_sim_factual + b200i_theta_gram: same outputs, same Gram, moments to rounding."""
    import torch
    params, draws = h.random_cohort(5000, seed=78)
    dmax = 12.999999999999998
    params['radio_sigmoid_betas'][900:940] = 6.0 / dmax          # some tiles take the generic fallback
    pd_ = dev.to_device(dev.pack_params(params))
    static = dev.to_device(np.asarray(params['patient_types'], dtype=np.float64))
    ins = []
    for k in ('noise', 'recovery', 'chemo', 'radio'):
        t = dev.alloc_rows(5000, 60, pitch)
        t.copy_(torch.from_numpy(draws[k]).cuda())
        ins.append(t)
    out_a, _ = dev.sim_factual(pd_, *ins, 60)
    ref = dev.theta_gram(out_a['cancer_volume'], out_a['chemo_application'], out_a['radio_application'],
                         out_a['sequence_lengths'], static, out_a['chemo_dosage'], out_a['radio_dosage'], tag="ta").clone()
    out_b, codes, pm = dev.sim_factual_side(pd_, *ins, 60)
    lean = dev.theta_gram_codes(out_b['cancer_volume'], codes, out_b['sequence_lengths'], static, pm, tag="tb").clone()
    torch.cuda.synchronize()
    for k in out_a:
        assert torch.equal(out_a[k], out_b[k]), k
    want = (out_a['chemo_application'] + 2 * out_a['radio_application']).to(torch.uint8)
    assert torch.equal(codes[:, :60], want)
    a, b = ref.cpu().numpy(), lean.cpu().numpy()
    assert np.array_equal(a[:60], b[:60])                        # Gram / right-hand sides / counts: same walk
    np.testing.assert_allclose(b[60:], a[60:], rtol=1e-12)       # moments: different summation order


def test_theta_gram_pitched_rows_bit_identical_to_dense(dev, seed1):
    """b200i_theta_gram_pitched: rows padded to 128-byte lines give the same statistics bit for bit."""
    import torch
    _, o, _, _ = seed1
    sim = o['train']
    dense = _device_stats(dev, sim, sim['patient_types']).clone()
    n, T = sim['cancer_volume'].shape

    def pitched(a):
        t = dev.alloc_rows(n, T, dev.aligned_pitch(T))
        t._base.fill_(float('nan'))
        t.copy_(torch.from_numpy(np.ascontiguousarray(a)).cuda())
        return t
    stats = dev.theta_gram(pitched(sim['cancer_volume']), pitched(sim['chemo_application']),
                           pitched(sim['radio_application']), dev.to_device(sim['sequence_lengths']),
                           dev.to_device(np.asarray(sim['patient_types'], dtype=np.float64)),
                           pitched(sim['chemo_dosage']), pitched(sim['radio_dosage']))
    torch.cuda.synchronize()
    assert torch.equal(stats, dense)


def test_population_stlsq_reproduces_reference_log(dev, seed1):
    import torch
    _, o, _, _ = seed1
    log = h.load_json('ref_log_seed1.json')['sindy']
    stats = _device_stats(dev, o['train'], o['train']['patient_types'])
    coefs, support = dev.stlsq_population(stats, threshold=1e-3, alpha=0.5, max_iter=100)
    torch.cuda.synchronize()
    assert support.cpu().numpy().all()
    np.testing.assert_allclose(coefs.cpu().numpy(), np.array(log['coefs']), rtol=1e-8)


def test_fused_simulator_statistics_equal_standalone(dev):
    import torch
    from oracle import sim_oracle as so
    params, draws = h.random_cohort(5000, seed=77)
    ref = so.sim_factual(params, 60, draws)
    pd_ = dev.to_device(dev.pack_params(params))
    static = dev.to_device(np.asarray(params['patient_types'], dtype=np.float64))
    args = [dev.to_device(draws[k]) for k in ('noise', 'recovery', 'chemo', 'radio')]
    for variant in (1, 2, 0, 10, 12):
        out, fused = dev.sim_factual(pd_, *args, 60, variant=variant, fused_static=static)
        fused = fused.clone()
        alone = dev.theta_gram(out['cancer_volume'], out['chemo_application'], out['radio_application'],
                               out['sequence_lengths'], static, out['chemo_dosage'], out['radio_dosage'], tag="t2")
        torch.cuda.synchronize()
        np.testing.assert_allclose(fused.cpu().numpy(), alone.cpu().numpy(), rtol=1e-12, atol=1e-9)
    G, b, cnt = _oracle_gram(ref)
    u = dev.unpack_stats(fused.cpu().numpy())
    assert np.array_equal(u['count'], cnt)
    np.testing.assert_allclose(u['G'], G, rtol=1e-10)
    # lean kernel with tiles that fall back to the generic column function (mixed sigmoids): same statistics
    dmax = 12.999999999999998
    params['radio_sigmoid_betas'][700:760] = 6.0 / dmax
    pd2 = dev.to_device(dev.pack_params(params))
    out, fused = dev.sim_factual(pd2, *args, 60, variant=12, fused_static=static)
    fused = fused.clone()
    alone = dev.theta_gram(out['cancer_volume'], out['chemo_application'], out['radio_application'],
                           out['sequence_lengths'], static, out['chemo_dosage'], out['radio_dosage'], tag="t2")
    torch.cuda.synchronize()
    np.testing.assert_allclose(fused.cpu().numpy(), alone.cpu().numpy(), rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("threshold,alpha", [(1e-3, 0.5), (0.08, 0.5), (0.6, 0.05), (5.0, 0.5)])
def test_population_stlsq_support_matches_oracle(dev, seed1, threshold, alpha):
    """Sparsity support bit-exact vs pysindy-semantics STLSQ at thresholds that prune 0..all terms."""
    import torch
    from oracle import sindy_np as sp
    _, o, means, stds = seed1
    dtr, sc = sp.process_data(o['train'], means, stds)
    buckets = sp.de_format_snippets(dtr, sc)
    stats = _device_stats(dev, o['train'], o['train']['patient_types'])
    coefs, support = dev.stlsq_population(stats, threshold=threshold, alpha=alpha)
    torch.cuda.synchronize()
    coefs, support = coefs.cpu().numpy(), support.cpu().numpy().astype(bool)
    import warnings
    warnings.filterwarnings('ignore')
    for a in range(4):
        th, xd = sp.design_matrices(buckets[a])
        c_ref, ind_ref = sp.stlsq_fit(th, xd, threshold, alpha)
        assert np.array_equal(support[a], ind_ref), (a, support[a], ind_ref)
        np.testing.assert_allclose(coefs[a], c_ref, rtol=1e-7, atol=1e-12)


def test_rollout_fp32_within_1e4_of_fp64(dev, seed1):
    """FP32 variant of the discovered-ODE rollout (BASELINE config C4): 1e-4 relative to the FP64 path."""
    import torch
    _, o, _, _ = seed1
    log = h.load_json('ref_log_seed1.json')['sindy']
    coefs = dev.to_device(np.array(log['coefs']))
    one = o['one']
    R, T = one['cancer_volume'].shape
    codes = dev.treatment_codes(dev.to_device(one['chemo_application']), dev.to_device(one['radio_application']), T - 1)
    x0 = dev.to_device(np.ascontiguousarray(one['cancer_volume'][:, 0]))
    static = dev.to_device(np.asarray(one['patient_types'], dtype=np.float64))
    p64 = dev.ode_rollout(x0, static, codes, coefs)
    p32 = dev.ode_rollout(x0, static, codes, coefs, fp32=True)
    torch.cuda.synchronize()
    a, b = p64.cpu().numpy(), p32.cpu().numpy()
    assert np.isfinite(b).all()
    scale = np.maximum(np.abs(a), 1e-3 * np.abs(a).max())       # relative to the trajectory's scale
    assert np.max(np.abs(a - b) / scale) < 1e-4                 # north star: 1e-4 in FP32
    assert not np.array_equal(a, b)                             # it really is a different arithmetic


def test_rollout_and_metrics_reproduce_reference_log(dev, seed1):
    """K6 + masked-SE on the one-step and treatment-sequence test sets: the 8 logged SINDy RMSEs."""
    import torch
    from oracle import sindy_np as sp
    _, o, means, stds = seed1
    log = h.load_json('ref_log_seed1.json')['sindy']
    coefs = dev.to_device(np.array(log['coefs']))
    norm = 1150.3465099894624
    # one-step set
    one = o['one']
    R, T = one['cancer_volume'].shape
    vol = dev.to_device(one['cancer_volume'])
    codes = dev.treatment_codes(dev.to_device(one['chemo_application']), dev.to_device(one['radio_application']), T - 1)
    pred = dev.ode_rollout(vol[:, 0].contiguous(), dev.to_device(one['patient_types']), codes, coefs)
    target = vol[:, 1:].contiguous()
    active = dev.to_device(one['sequence_lengths'], dtype=torch.int32)
    sums = dev.masked_se(pred, target, active).cpu().numpy()
    W = T - 1
    se, cnt, tl, cl = sums[:W], sums[W:2 * W], sums[3 * W], sums[3 * W + 1]
    rmse_all = np.sqrt(se.sum() / cnt.sum()) / norm * 100
    rmse_orig = np.sqrt((se / cnt).mean()) / norm * 100
    rmse_last = np.sqrt(tl / cl) / norm * 100
    np.testing.assert_allclose([rmse_all, rmse_orig, rmse_last],
                               [log['encoder_test_rmse_all'], log['encoder_test_rmse_orig'],
                                log['encoder_test_rmse_last']], rtol=1e-9)
    # element-wise against the oracle rollout
    d1, sc = sp.process_data(one, means, stds)
    ref_pred = sp.predictions_population(d1, sc, np.array(log['coefs']))[..., 0] * sc['output_stds'] + sc['output_means']
    np.testing.assert_allclose(pred.cpu().numpy(), ref_pred, rtol=1e-9, atol=1e-9)
    # treatment sequences: 5 predictions at [max(1, sl-5), +5)
    seq = o['seq']
    R, Wf = seq['cancer_volume'].shape
    vol = dev.to_device(seq['cancer_volume'])
    codes = dev.treatment_codes(dev.to_device(seq['chemo_application']), dev.to_device(seq['radio_application']), Wf - 1)
    pred = dev.ode_rollout(vol[:, 0].contiguous(), dev.to_device(seq['patient_types']), codes, coefs)
    sl = seq['sequence_lengths'].astype(np.int64)
    lo = np.clip(np.maximum(1, sl - 5), 0, Wf - 1 - 5)
    idx = torch.from_numpy(lo[:, None] + np.arange(5)[None, :]).cuda()
    p5 = torch.gather(pred, 1, idx).contiguous()
    # targets: process_sequential_test takes outputs[fact_length : fact_length+5], fact_length = sl-5
    t5 = torch.gather(vol[:, 1:], 1, torch.from_numpy((sl - 5)[:, None] + np.arange(5)[None, :]).cuda()).contiguous()
    sums = dev.masked_se(p5, t5, torch.full((R,), 5, dtype=torch.int32, device='cuda')).cpu().numpy()
    rm = np.sqrt(sums[:5] / sums[5:10]) / norm * 100
    np.testing.assert_allclose(rm, log['decoder_test_rmse_2_to_6_step'], rtol=1e-9)
