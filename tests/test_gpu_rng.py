"""K1L (device-generated draws): the Philox generator vs its numpy restatement, and the fused simulator vs the
pre-drawn-array contract of K1 on the same draws.  All calls go through the C ABI (b200_insite.device)."""
import numpy as np
import pytest

import helpers as h

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from b200_insite import device as dev
    dev.require_cuda()
    return dev


def test_philox_draws_match_numpy_restatement(dev):
    """Uniforms bit-exact; Box-Muller noise to 1e-12 relative (1e-16 absolute near the zeros of cos/sin) (libdevice vs libm log / sincos, <= 2 ulp each)."""
    import torch
    from oracle import philox_np as ph
    for n, T, seed, base, pitch in ((257, 60, 7, 0, 60), (100, 62, (5 << 32) + 11, (1 << 33) + 5, 64)):
        got = dev.philox_draws(n, T, seed, patient_base=base, pitch=pitch)
        torch.cuda.synchronize()
        want = ph.draw_factual(n, T, seed, patient_base=base)
        for g, k in zip(got[1:], ('recovery', 'chemo', 'radio')):
            assert np.array_equal(g.cpu().numpy(), want[k]), k
        np.testing.assert_allclose(got[0].cpu().numpy(), want["noise"], rtol=1e-12, atol=1e-16)   # cos(pi x) near its zeros
    u = got[2].cpu().numpy()
    assert u.min() >= 0.0 and u.max() < 1.0


def _cohort(n, seed, slow_rows=None):
    params, _ = h.random_cohort(n, seed=seed)
    if slow_rows is not None:   # mixed sigmoids: those 32-patient tiles take the generic column function
        params['radio_sigmoid_betas'][slow_rows] = 6.0 / 12.999999999999998
    return params


@pytest.mark.parametrize("n,T,pitch,slow", [(5000, 60, 64, None), (4999, 60, 60, slice(900, 940)), (333, 62, 64, None),
                                            (70, 20, 20, slice(0, 3)), (31, 4, 4, None)])
def test_sim_factual_rng_equals_k1_on_exported_draws(dev, n, T, pitch, slow):
    """b200i_sim_factual_rng == b200i_sim_factual fed with b200i_philox_draws, bit for bit: volumes, treatment codes,
    sequence lengths, per-patient moment sums."""
    import torch
    params = _cohort(n, 81, slow)
    pd_ = dev.to_device(dev.pack_params(params))
    seed, base = 20231018, 12345
    draws = dev.philox_draws(n, T, seed, patient_base=base, pitch=pitch)
    out, _ = dev.sim_factual(pd_, *draws, T)
    vol, codes, sl, pm, _ = dev.sim_factual_rng(pd_, T, seed, patient_base=base, pitch=pitch)
    g1 = dev.sim_factual_rng(pd_, T, seed, patient_base=base, pitch=pitch, variant=1)   # first-generation kernel
    torch.cuda.synchronize()
    for a, b in zip((vol, codes, sl, pm), g1):
        assert torch.equal(a, b)
    assert torch.equal(vol, out['cancer_volume'])
    assert torch.equal(sl, out['sequence_lengths'])
    want = (out['chemo_application'] + 2 * out['radio_application']).to(torch.uint8)
    assert torch.equal(codes[:, :T], want)
    assert int(codes[:, T:].sum()) == 0
    # moments: sums over the active entries (inactive entries are zero)
    v, c, d = out['cancer_volume'].cpu().numpy(), out['chemo_dosage'].cpu().numpy(), out['radio_dosage'].cpu().numpy()
    got = pm.cpu().numpy()
    for j, a in enumerate((v, v * v, c, c * c, d, d * d)):
        np.testing.assert_allclose(got[j], a.sum(axis=1), rtol=1e-12, atol=1e-300)
    assert 2 < float(sl.mean()) <= T - 1


def test_sim_factual_rng_is_shard_invariant(dev):
    """Counter = global patient index: two shards with patient_base reproduce the single launch bit for bit."""
    import torch
    n, T = 3000, 60
    params = _cohort(n, 82)
    block = dev.pack_params(params)
    whole = dev.sim_factual_rng(dev.to_device(block), T, 99)
    parts = [dev.sim_factual_rng(dev.to_device(block[:, a:b]), T, 99, patient_base=a) for a, b in ((0, 1111), (1111, n))]
    torch.cuda.synchronize()
    for j in range(4):
        cat = torch.cat([p[j] for p in parts], dim=1 if j == 3 else 0)
        assert torch.equal(cat, whole[j]), j


@pytest.mark.parametrize("slow", [None, slice(700, 760)])
def test_sim_factual_rng_fused_statistics(dev, slow):
    """Fused population statistics of the generator kernel == theta_gram on K1's arrays for the same draws, and the
    lean pair (moments + codes -> theta_gram_codes) gives the same; STLSQ support identical."""
    import torch
    n, T = 5000, 60
    params = _cohort(n, 83, slow)
    pd_ = dev.to_device(dev.pack_params(params))
    static = dev.to_device(np.asarray(params['patient_types'], dtype=np.float64))
    draws = dev.philox_draws(n, T, 5, pitch=64)
    out, _ = dev.sim_factual(pd_, *draws, T)
    alone = dev.theta_gram(out['cancer_volume'], out['chemo_application'], out['radio_application'],
                           out['sequence_lengths'], static, out['chemo_dosage'], out['radio_dosage'], tag="r0").clone()
    vol, codes, sl, pm, fused = dev.sim_factual_rng(pd_, T, 5, fused_static=static, tag="r1")
    fused = fused.clone()
    fused1 = dev.sim_factual_rng(pd_, T, 5, fused_static=static, tag="r3", variant=1)[4].clone()
    vol2, codes2, sl2, pm2, _ = dev.sim_factual_rng(pd_, T, 5)
    lean = dev.theta_gram_codes(vol2, codes2, sl2, static, pm2, tag="r2").clone()
    torch.cuda.synchronize()
    assert torch.equal(vol, vol2) and torch.equal(codes, codes2)
    a, f, l = alone.cpu().numpy(), fused.cpu().numpy(), lean.cpu().numpy()
    np.testing.assert_allclose(f, a, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(fused1.cpu().numpy(), a, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(l, a, rtol=1e-12, atol=1e-9)
    assert np.array_equal(dev.unpack_stats(f)['count'], dev.unpack_stats(a)['count'])
    ca, sa = dev.stlsq_population(alone)
    cf, sf = dev.stlsq_population(fused)
    torch.cuda.synchronize()
    assert torch.equal(sa, sf)
    np.testing.assert_allclose(cf.cpu().numpy(), ca.cpu().numpy(), rtol=1e-9, atol=1e-13)


def test_sim_factual_rng_argument_errors(dev):
    import torch
    params = _cohort(64, 84)
    pd_ = dev.to_device(dev.pack_params(params))
    with pytest.raises(RuntimeError, match="seq_length"):
        dev.sim_factual_rng(pd_, 61, 1, volume=dev.alloc_rows(64, 61, 62))
    with pytest.raises(RuntimeError, match="code_pitch"):
        dev.sim_factual_rng(pd_, 60, 1, codes=torch.empty((64, 60), dtype=torch.uint8, device='cuda'))


@pytest.mark.parametrize("chunks", [1, 3, 8])
def test_generated_pipeline_host_step_equals_device_step(dev, chunks):
    """GeneratedFitPipeline.step_host (pinned host parameters, chunked H2D overlapped with K1L through
    b200i_upload_simulate_rng) == step_device on resident parameters: cohort and statistics bit for bit."""
    import torch
    from b200_insite.cohort import GeneratedFitPipeline
    n, T = 4133, 60
    params = _cohort(n, 85)
    block = torch.from_numpy(dev.pack_params(params)).pin_memory()
    static = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.float64)).pin_memory()
    a = GeneratedFitPipeline(n, T, seed=3, patient_base=1000, chunks=chunks)
    b = GeneratedFitPipeline(n, T, seed=3, patient_base=1000)
    b.params.copy_(block); b.static.copy_(static)
    res = torch.zeros(32 + dev.STATS_DOUBLES, dtype=torch.float64).pin_memory()
    a.step_host(block, static, res)
    coefs = b.step_device()
    torch.cuda.synchronize()
    assert a.chunks == len(a.bounds) <= chunks
    assert torch.equal(a.volume, b.volume) and torch.equal(a.codes, b.codes)
    assert torch.equal(a.sequence_lengths, b.sequence_lengths) and torch.equal(a.patient_moments, b.patient_moments)
    # statistics: per-chunk shares summed in chunk order vs one launch (same values, different summation order)
    np.testing.assert_allclose(a.stats.cpu().numpy(), b.stats.cpu().numpy(), rtol=1e-12, atol=1e-9)
    assert np.array_equal(a.stats.cpu().numpy()[[14, 29, 44, 59, 66, 67]], b.stats.cpu().numpy()[[14, 29, 44, 59, 66, 67]])  # counts
    assert np.array_equal(res[16:32].numpy().reshape(4, 4) != 0, b.support.cpu().numpy() != 0)
    np.testing.assert_allclose(res[:16].numpy().reshape(4, 4), coefs.cpu().numpy(), rtol=1e-9, atol=1e-13)
    assert np.array_equal(res[32:].numpy(), a.stats.cpu().numpy())
    if chunks == 1:
        assert torch.equal(a.stats, b.stats)
    # parameter rows that are one scalar for the cohort (K and the four sigmoid rows of generate_params) are filled on the
    # device instead of being copied: same cohort, and a host block whose uniform rows hold garbage is never read there
    uni = dev.uniform_param_rows(params)
    assert sorted(uni) == [5, 6, 7, 8, 9]
    junk = block.clone()
    junk[5:10] = float('nan')
    c = GeneratedFitPipeline(n, T, seed=3, patient_base=1000, chunks=chunks)
    c.params.fill_(-1.0)
    c.step_host(junk.pin_memory(), static, res, uniform=uni)
    torch.cuda.synchronize()
    assert torch.equal(c.params, b.params) and torch.equal(c.volume, b.volume) and torch.equal(c.codes, b.codes)
    assert torch.equal(c.stats, a.stats) and c.h2d_bytes(uni) == 6 * n * 8
    # reduced input set (what get_standard_params draws): beta = alpha / 10 and the static feature are rebuilt on the
    # device, the scalar rows come from generate_params' arguments -- 33 bytes per patient, the same bits
    assert dev.cohort_scalar_rows(2.0, 2.0) == uni
    assert np.array_equal(params['beta'], params['alpha'] / 10)
    junk[3] = float('nan')
    types = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.uint8)).pin_memory()
    d = GeneratedFitPipeline(n, T, seed=3, patient_base=1000, chunks=chunks)
    d.params.fill_(-1.0); d.static.fill_(-1.0)
    d.step_host(junk.pin_memory(), None, res, uniform=dev.cohort_scalar_rows(2.0, 2.0), types_u8=types)
    torch.cuda.synchronize()
    assert torch.equal(d.params, b.params) and torch.equal(d.static, b.static)
    assert torch.equal(d.volume, b.volume) and torch.equal(d.codes, b.codes) and torch.equal(d.stats, a.stats)
    assert d.h2d_bytes(uni, reduced=True) == 33 * n
    # the same step captured into a CUDA graph and replayed: new CONTENT in the same pinned buffers, same bits as eager
    hp = junk.pin_memory()
    e = GeneratedFitPipeline(n, T, seed=3, patient_base=1000, chunks=chunks)
    res2 = torch.zeros_like(res).pin_memory()
    for rep in range(3):
        if rep == 2:      # refresh the inputs between replays
            params2 = _cohort(n, 86)
            blk2 = torch.from_numpy(dev.pack_params(params2))
            hp.copy_(blk2); types.copy_(torch.from_numpy(np.asarray(params2['patient_types'], dtype=np.uint8)))
        e.step_host(hp, None, res2, uniform=dev.cohort_scalar_rows(2.0, 2.0), types_u8=types, graph=True)
    f = GeneratedFitPipeline(n, T, seed=3, patient_base=1000, chunks=chunks)
    res3 = torch.zeros_like(res).pin_memory()
    f.step_host(hp, None, res3, uniform=dev.cohort_scalar_rows(2.0, 2.0), types_u8=types)
    torch.cuda.synchronize()
    assert torch.equal(e.volume, f.volume) and torch.equal(e.codes, f.codes) and torch.equal(e.stats, f.stats)
    assert torch.equal(res2, res3) and not torch.equal(e.volume, d.volume)


@pytest.mark.parametrize("chunks", [3, 8])
def test_submitted_steps_overlap_and_equal_the_serial_steps(dev, chunks):
    """GeneratedFitPipeline.submit: a stream of cohorts whose uploads run under the previous step's tail.  Two different
    cohorts alternate through two sets of pinned buffers; every result equals the serial step_host result of its cohort
    bit for bit (a copy that overtook a reader of the previous step would show up here)."""
    import torch
    from b200_insite.cohort import GeneratedFitPipeline
    n, T = 50021, 60
    uni = dev.cohort_scalar_rows(2.0, 2.0)
    sets, want = [], []
    for seed in (85, 86):
        params = _cohort(n, seed)
        hp = torch.from_numpy(dev.pack_params(params)).pin_memory()
        ty = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.uint8)).pin_memory()
        ref = GeneratedFitPipeline(n, T, seed=3, patient_base=7, chunks=chunks)
        r = torch.zeros(32 + dev.STATS_DOUBLES, dtype=torch.float64).pin_memory()
        ref.step_host(hp, None, r, uniform=uni, types_u8=ty)
        sets.append((hp, ty)); want.append(r.clone())
    assert not torch.equal(want[0], want[1])
    pipe = GeneratedFitPipeline(n, T, seed=3, patient_base=7, chunks=chunks)
    results = [torch.zeros(32 + dev.STATS_DOUBLES, dtype=torch.float64).pin_memory() for _ in range(2)]
    pending, got = None, []
    for s in range(7):
        k = s & 1
        step = pipe.submit(sets[k][0], results[k], sets[k][1], uniform=uni)
        if pending is not None:
            got.append((pending[0], pending[1].wait().clone()))
        pending = (k, step)
    got.append((pending[0], pending[1].wait().clone()))
    pending[1].inputs_consumed.synchronize()
    assert len(got) == 7
    for k, r in got:
        assert torch.equal(r, want[k])
    # and a serial step after the stream still sees a quiescent pipeline
    r = torch.zeros_like(results[0]).pin_memory()
    pipe.step_host(sets[1][0], None, r, uniform=uni, types_u8=sets[1][1])
    assert torch.equal(r, want[1])


def test_generated_counterfactual_draws_and_cohort(dev):
    """counterfactual.generated_draws: the device generator's draws in the layout of the counterfactual simulators
    (noise (N, T+H) with odd width) equal the numpy restatement, and K3 on them equals the C oracle."""
    import torch
    from oracle import philox_np as ph, sim_oracle as so
    from b200_insite import counterfactual as cfm
    n, T, H = 40, 60, 5
    noise, rec, chemo, radio = cfm.generated_draws(n, T, H, seed=21, patient_base=7)
    torch.cuda.synchronize()
    want = ph.draw_factual(n, T + H + 1, 21, patient_base=7)
    assert noise.shape == (n, T + H) and rec.shape == (n, T)
    np.testing.assert_allclose(noise.cpu().numpy(), want['noise'][:, :T + H], rtol=1e-12, atol=1e-16)
    for g, k in ((rec, 'recovery'), (chemo, 'chemo'), (radio, 'radio')):
        assert np.array_equal(g.cpu().numpy(), want[k][:, :T])
    params = _cohort(n, 86)
    draws = {'noise': noise.cpu().numpy(), 'recovery': rec.cpu().numpy(), 'chemo': chemo.cpu().numpy(),
             'radio': radio.cpu().numpy()}
    ref = so.sim_cf_treatment_seq(params, T, H, draws)
    coh = cfm.sim_cf_treatment_seq(dev.to_device(dev.pack_params(params)), noise, rec, chemo, radio, T, H)
    dense = cfm.expand(coh, dev.to_device(np.asarray(params['patient_types'], dtype=np.float64)))
    torch.cuda.synchronize()
    assert coh.total_rows == ref['cancer_volume'].shape[0]
    for k in ('chemo_application', 'radio_application', 'sequence_lengths', 'patient_current_t'):
        assert np.array_equal(dense[k].cpu().numpy(), ref[k]), k
    np.testing.assert_allclose(dense['cancer_volume'].cpu().numpy(), ref['cancer_volume'], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("n", [0, 1, 33])
def test_sim_factual_rng_tiny_and_empty_cohorts(dev, n):
    """Empty, single-patient and ragged (one full tile + one patient) cohorts: same results as K1 on the exported
    draws; the empty cohort leaves zero statistics."""
    import torch
    T = 60
    params = _cohort(max(n, 1), 87)
    block = dev.pack_params(params)[:, :n]
    pd_ = dev.to_device(block) if n else torch.empty((10, 0), dtype=torch.float64, device='cuda')
    static = dev.to_device(np.asarray(params['patient_types'], dtype=np.float64)[:n]) if n else \
        torch.empty((0,), dtype=torch.float64, device='cuda')
    vol, codes, sl, pm, _ = dev.sim_factual_rng(pd_, T, 5, pitch=T)
    stats = dev.theta_gram_codes(vol, codes, sl, static, pm, tag="tiny").clone()
    torch.cuda.synchronize()
    assert vol.shape == (n, T) and codes.shape[0] == n and sl.shape == (n,)
    if n == 0:
        assert float(stats.abs().sum()) == 0.0
        return
    draws = dev.philox_draws(n, T, 5, pitch=T)
    out, _ = dev.sim_factual(pd_, *draws, T)
    alone = dev.theta_gram(out['cancer_volume'], out['chemo_application'], out['radio_application'],
                           out['sequence_lengths'], static, out['chemo_dosage'], out['radio_dosage'], tag="tiny2")
    torch.cuda.synchronize()
    assert torch.equal(vol, out['cancer_volume']) and torch.equal(sl, out['sequence_lengths'])
    np.testing.assert_allclose(stats.cpu().numpy(), alone.cpu().numpy(), rtol=1e-12, atol=1e-9)


def test_sim_factual_rng_full_size_equals_k1_on_exported_draws(dev):
    """BASELINE size (1M patients x 60 steps): the generated-draws kernel and K1 on the exported draws agree bit for
    bit (row indices beyond 2^16 tiles, 64-bit offsets), and the chunked host pipeline reproduces the same cohort."""
    import torch
    from b200_insite.cohort import GeneratedFitPipeline
    n, T, seed, base = 1_000_000, 60, 77, 3_000_000_000
    params = _cohort(n, 88)
    block = torch.from_numpy(dev.pack_params(params))
    pd_ = block.cuda()
    draws = dev.philox_draws(n, T, seed, patient_base=base, pitch=64)
    out, _ = dev.sim_factual(pd_, *draws, T)
    del draws
    vol, codes, sl, pm, _ = dev.sim_factual_rng(pd_, T, seed, patient_base=base)
    torch.cuda.synchronize()
    assert torch.equal(vol, out['cancer_volume']) and torch.equal(sl, out['sequence_lengths'])
    assert torch.equal(codes[:, :T], (out['chemo_application'] + 2 * out['radio_application']).to(torch.uint8))
    del out
    pipe = GeneratedFitPipeline(n, T, seed=seed, patient_base=base, chunks=16)
    static = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.float64)).pin_memory()
    res = torch.zeros(32 + dev.STATS_DOUBLES, dtype=torch.float64).pin_memory()
    pipe.step_host(block.pin_memory(), static, res)
    torch.cuda.synchronize()
    assert torch.equal(pipe.volume, vol) and torch.equal(pipe.codes, codes) and torch.equal(pipe.patient_moments, pm)
    assert 55.0 < float(sl.mean()) < 58.0 and np.isfinite(res.numpy()).all()
