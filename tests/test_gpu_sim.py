"""GPU parity tests of the simulators (K1, K2, K3) against the oracle and the reference fixtures.

Tolerances (BASELINE.json north_star): treatment assignments, flags, sequence lengths and row counts
bit-exact; float64 trajectories within 1e-6 relative -- asserted here at 1e-9 (measured ~1e-13)."""
import numpy as np
import pytest

import helpers as h

pytestmark = pytest.mark.gpu

FLOAT_RTOL = 1e-9
EXACT_F = ('chemo_application', 'radio_application', 'radio_dosage', 'death_flags', 'recovery_flags',
           'sequence_lengths')
FLOAT_F = ('cancer_volume', 'chemo_dosage', 'chemo_probabilities', 'radio_probabilities')


@pytest.fixture(scope="module")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from b200_insite import device
    device.require_cuda()
    return device


def _run_factual(dev, params, draws, T=60, variant=0, assigned=None, fused=False):
    import torch
    pd_ = dev.to_device(dev.pack_params(params))
    static = dev.to_device(np.asarray(params['patient_types'], dtype=np.float64)) if fused else None
    aa = None if assigned is None else dev.to_device(assigned)
    consts = dev.sim_consts(params['window_size'], params['lag'])
    out, stats = dev.sim_factual(pd_, dev.to_device(draws['noise']), dev.to_device(draws['recovery']),
                                 dev.to_device(draws['chemo']), dev.to_device(draws['radio']), T, consts,
                                 assigned_actions=aa, variant=variant, fused_static=static)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}, (stats.cpu().numpy() if stats is not None else None)


def _check_factual(got, ref, tag):
    for k in EXACT_F:
        assert np.array_equal(got[k], ref[k]), f"{tag}: {k} not bit-exact ({np.count_nonzero(got[k] != ref[k])} diffs)"
    for k in FLOAT_F:
        np.testing.assert_allclose(got[k], ref[k], rtol=FLOAT_RTOL, atol=1e-13, err_msg=f"{tag}: {k}")


@pytest.mark.parametrize("variant", [1, 2, 10, 12])
def test_factual_variants_match_oracle(dev, variant):
    from oracle import sim_oracle as so
    params, draws = h.random_cohort(3000, seed=5)   # ragged: last tile partial for every tile size
    ref = so.sim_factual(params, 60, draws)
    got, _ = _run_factual(dev, params, draws, variant=variant)
    _check_factual(got, ref, f"variant {variant}")


@pytest.mark.parametrize("variant", [0, 10, 12])
def test_factual_fast_path_preconditions_fall_back_per_tile(dev, variant):
    """Tiles whose patients break the lean kernel's preconditions (chemo and radio sigmoids differ, sigmoid
    argument beyond exp_fast's domain) are processed by the generic column function inside the same launch."""
    from oracle import sim_oracle as so
    params, draws = h.random_cohort(1000, seed=21)
    dmax = 12.999999999999998
    params['radio_sigmoid_betas'][100:140] = 6.0 / dmax          # differs from the chemo sigmoid
    params['chemo_sigmoid_intercepts'][300:303] = 5.0
    params['chemo_sigmoid_betas'][500:520] = 5000.0 / dmax        # |z| up to ~2500: exp overflows to inf / 0
    params['radio_sigmoid_betas'][500:520] = 5000.0 / dmax
    ref = so.sim_factual(params, 60, draws)
    got, _ = _run_factual(dev, params, draws, variant=variant)
    _check_factual(got, ref, f"fallback variant {variant}")


@pytest.mark.parametrize("variant,pitch", [(0, 64), (10, 64), (12, 64), (0, 62), (10, 72)])
def test_factual_pitched_rows_bit_identical_to_dense(dev, variant, pitch):
    """cudaMallocPitch-style rows (b200i_sim_factual_pitched): same kernels, same bits, padding untouched."""
    import torch
    from oracle import sim_oracle as so
    params, draws = h.random_cohort(1500, seed=9)
    ref = so.sim_factual(params, 60, draws)
    pd_ = dev.to_device(dev.pack_params(params))
    ins = []
    for k in ('noise', 'recovery', 'chemo', 'radio'):
        t = dev.alloc_rows(1500, 60, pitch)
        t.copy_(torch.from_numpy(draws[k]).cuda())
        ins.append(t)
    out = {k: dev.alloc_rows(1500, 60, pitch) for k in dev.FACTUAL_OUT_KEYS}
    for k in out:
        out[k]._base.fill_(-7.0)        # padding columns must survive
    out['sequence_lengths'] = torch.empty((1500,), dtype=torch.float64, device='cuda')
    got, _ = dev.sim_factual(pd_, *ins, 60, out=out, variant=variant)
    torch.cuda.synchronize()
    for k in dev.FACTUAL_OUT_KEYS:
        assert bool((got[k]._base[:, 60:] == -7.0).all()), f"{k}: padding overwritten"
    _check_factual({k: v.cpu().numpy() for k, v in got.items()}, ref, f"pitch {pitch} variant {variant}")


def test_factual_matches_reference_fixture(dev):
    """Reference outputs themselves (tests/golden/ref_sim_small.npz), reference RNG order."""
    g = h.load_npz('ref_sim_small.npz')
    inputs = h.collection_inputs(7, 2.0, 192, 24, 24)
    for name in ('train', 'val'):
        got, _ = _run_factual(dev, *inputs[name])
        ref = {k: g[f'{name}/out/{k}'] for k in EXACT_F + FLOAT_F}
        _check_factual(got, ref, name)


def test_factual_dropin_consumes_global_rng_like_reference(dev):
    import b200_insite.cancer_simulation as cs
    g = h.load_npz('ref_sim_small.npz')
    np.random.seed(7)
    p = cs.generate_params(192, 2.0, 2.0, 15, 0)
    out = cs.simulate_factual(p, 60)
    assert set(out) == {'cancer_volume', 'chemo_dosage', 'radio_dosage', 'chemo_application', 'radio_application',
                        'chemo_probabilities', 'radio_probabilities', 'sequence_lengths', 'death_flags',
                        'recovery_flags', 'patient_types'}
    _check_factual(out, {k: g[f'train/out/{k}'] for k in EXACT_F + FLOAT_F}, 'drop-in train')
    p2 = cs.generate_params(24, 2.0, 2.0, 15, 0)     # RNG stream continues: validation subset matches too
    out2 = cs.simulate_factual(p2, 60)
    _check_factual(out2, {k: g[f'val/out/{k}'] for k in EXACT_F + FLOAT_F}, 'drop-in val')
    m, s = cs.get_scaling_params(out)
    np.testing.assert_allclose(m.values, g['train/scaling_means'], rtol=1e-12)
    np.testing.assert_allclose(s.values, g['train/scaling_stds'], rtol=1e-12)


def test_factual_gamma10_and_assigned_actions_fixture(dev):
    from oracle import rng_export as rx
    g = h.load_npz('ref_sim_gamma10.npz')
    np.random.seed(100)
    p = rx.generate_params(128, 10.0, 10.0, 15, 0)
    d = rx.draw_factual(128, 60)
    got, _ = _run_factual(dev, p, d)
    _check_factual(got, {k: g[f'out/{k}'] for k in EXACT_F + FLOAT_F}, 'gamma10')
    d2 = rx.draw_factual(128, 60)
    got2, _ = _run_factual(dev, p, d2, assigned=g['assigned_actions'])
    _check_factual(got2, {k: g[f'out_assigned/{k}'] for k in EXACT_F + FLOAT_F}, 'assigned_actions')


@pytest.mark.parametrize("n,T,window", [(1, 60, 15), (127, 60, 15), (129, 60, 15), (700, 30, 15), (333, 61, 15),
                                        (500, 60, 7), (64, 4, 15)])
def test_factual_edge_shapes(dev, n, T, window):
    """Single patient, tile boundaries, short / odd horizons (odd T -> generic kernel), short window."""
    from oracle import sim_oracle as so
    params, draws = h.random_cohort(n, seed=n + T, T=T)
    params['window_size'] = window
    ref = so.sim_factual(params, T, draws)
    got, _ = _run_factual(dev, params, draws, T=T)
    _check_factual(got, ref, f"n={n} T={T} w={window}")


def test_factual_empty_cohort_and_bad_config(dev):
    import torch
    params, draws = h.random_cohort(4, seed=1)
    for k in h.PARAM_KEYS:
        params[k] = params[k][:0]
    draws = {k: v[:0] for k, v in draws.items()}
    got, _ = _run_factual(dev, params, draws)
    assert got['cancer_volume'].shape == (0, 60)
    params, draws = h.random_cohort(4, seed=1)
    params['lag'] = 1
    with pytest.raises(RuntimeError, match="lag"):
        _run_factual(dev, params, draws)


def test_factual_large_cohort_properties(dev):
    """1M patients (BASELINE config C2 size): oracle parity on a strided sample + structural properties."""
    import torch
    from oracle import sim_oracle as so
    n, T = 1_000_000, 60
    params, draws = h.random_cohort(n, seed=2024)
    got, _ = _run_factual(dev, params, draws)
    sl = got['sequence_lengths'].astype(np.int64)
    cols = np.arange(T)[None, :]
    assert sl.min() >= 2 and sl.max() == T - 1
    # nothing after the last simulated step, treatment never at t = 0, flags only at the last step
    for k in ('cancer_volume', 'chemo_dosage', 'radio_dosage', 'chemo_application', 'radio_application',
              'chemo_probabilities', 'radio_probabilities'):
        assert not np.any(got[k][cols >= sl[:, None]]), k
    assert not got['chemo_application'][:, 0].any() and not got['radio_application'][:, 0].any()
    assert np.array_equal(got['radio_dosage'], 2.0 * got['radio_application'])
    ended = got['death_flags'].sum(1) + got['recovery_flags'].sum(1)
    assert np.array_equal(ended > 0, sl < T - 1) or np.all((ended > 0) <= (sl <= T - 1))
    assert got['cancer_volume'].max() <= 1150.3465099894624
    # chemo dosage recursion C[t] = C[t-1]/2 + 5*app[t] holds exactly
    C, A = got['chemo_dosage'], got['chemo_application']
    act = cols[:, 1:] < sl[:, None]
    assert np.array_equal((C[:, 1:] == C[:, :-1] * 0.5 + 5.0 * A[:, 1:]) | ~act, np.ones_like(act))
    # oracle on every 97th patient
    idx = np.arange(0, n, 97)
    sub_p = {k: (v[idx] if isinstance(v, np.ndarray) else v) for k, v in params.items()}
    sub_d = {k: v[idx] for k, v in draws.items()}
    ref = so.sim_factual(sub_p, T, sub_d)
    _check_factual({k: v[idx] for k, v in got.items()}, ref, "1M sample")


# ---------------------------------------------------------------------------------------------------
# counterfactual generators
# ---------------------------------------------------------------------------------------------------
CF1_EXACT = ('chemo_application', 'radio_application', 'sequence_lengths', 'patient_types')
CFS_EXACT = CF1_EXACT + ('patient_ids_all_trajectories', 'patient_current_t')


def _run_cf(dev, kind, params, draws, T=60, H=5):
    from b200_insite import counterfactual as cf
    d = (draws['noise'], draws['recovery'], draws['chemo'], draws['radio'])
    if kind == 'one':
        return cf.one_step_dense(params, T, d)
    return cf.treatment_seq_dense(params, T, H, d)


def _check_cf(got, ref, exact, tag):
    assert got['cancer_volume'].shape == ref['cancer_volume'].shape, f"{tag}: row count"
    for k in exact:
        assert np.array_equal(got[k], ref[k]), f"{tag}: {k} not bit-exact"
    np.testing.assert_allclose(got['cancer_volume'], ref['cancer_volume'], rtol=FLOAT_RTOL, atol=1e-12,
                               err_msg=f"{tag}: cancer_volume")


def test_cf_generators_match_reference_fixture(dev):
    g = h.load_npz('ref_sim_small.npz')
    inputs = h.collection_inputs(7, 2.0, 192, 24, 24)
    got = _run_cf(dev, 'one', *inputs['one'])
    _check_cf(got, {k: g[f'one/out/{k}'] for k in CF1_EXACT + ('cancer_volume',)}, CF1_EXACT, 'one-step fixture')
    got = _run_cf(dev, 'seq', *inputs['seq'])
    _check_cf(got, {k: g[f'seq/out/{k}'] for k in CFS_EXACT + ('cancer_volume',)}, CFS_EXACT, 'treatment-seq fixture')


@pytest.mark.parametrize("n,seed", [(1, 3), (2, 4), (100, 1), (400, 9), (1500, 21)])
def test_cf_one_step_matches_oracle_across_wavefront_levels(dev, n, seed):
    from oracle import sim_oracle as so
    params, draws = h.random_cohort(n, seed=seed)
    ref = so.sim_cf_one_step(params, 60, draws)
    got = _run_cf(dev, 'one', params, draws)
    _check_cf(got, ref, CF1_EXACT, f"one-step n={n}")


@pytest.mark.parametrize("n,seed", [(1, 3), (3, 4), (100, 1), (700, 9)])
def test_cf_treatment_seq_matches_oracle_across_wavefront_levels(dev, n, seed):
    from oracle import sim_oracle as so
    params, draws = h.random_cohort(n, seed=seed, extra=5)
    ref = so.sim_cf_treatment_seq(params, 60, 5, draws)
    got = _run_cf(dev, 'seq', params, draws)
    _check_cf(got, ref, CFS_EXACT, f"treatment-seq n={n}")


def test_cf_dropin_entry_points_seed1_digests(dev):
    """Log configuration through the drop-in entry points (global RNG): row counts and the bit-exact
    arrays must reproduce the sha256 digests of the reference outputs."""
    import hashlib
    import b200_insite.cancer_simulation as cs
    dig = h.load_json('ref_digests_seed1.json')
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    np.random.seed(1)
    train = cs.simulate_factual(cs.generate_params(1000, 2.0, 2.0, 15, 0), 60)
    val = cs.simulate_factual(cs.generate_params(100, 2.0, 2.0, 15, 0), 60)
    one = cs.simulate_counterfactual_1_step(cs.generate_params(100, 2.0, 2.0, 15, 0), 60)
    seq = cs.simulate_counterfactuals_treatment_seq(cs.generate_params(100, 2.0, 2.0, 15, 0), 60, 5)
    for name, d in (('train', train), ('val', val), ('one', one), ('seq', seq)):
        assert d['cancer_volume'].shape[0] == dig[name]['rows'], name
        for k in ('chemo_application', 'radio_application', 'sequence_lengths'):
            assert sha(d[k]) == dig[name]['out_sha256'][k], (name, k)
        assert abs(d['cancer_volume'].sum() - dig[name]['cancer_volume_sum']) <= 1e-9 * abs(dig[name]['cancer_volume_sum'])
    with pytest.raises(NotImplementedError):
        cs.simulate_counterfactuals_treatment_seq(cs.generate_params(4, 2.0, 2.0, 15, 0), 60, 5, 'random_trajectories')


# ---- BASELINE-size counterfactual cohorts: wavefront depth 4, row indices past 2^31 / width -------------------------
def _sub_params(params, idx):
    return {k: (v[idx] if isinstance(v, np.ndarray) else v) for k, v in params.items()}


def _cohort_levels(off, n):
    """Dependency level of every patient from the row offsets: level(i) = 1 + level(owner of row i), patient 0 = 1."""
    owner = np.searchsorted(off, np.arange(n), side='right') - 1
    lvl = np.ones(n, dtype=np.int32)
    for i in range(1, n):                      # owner[i] < i
        lvl[i] = lvl[owner[i]] + 1
    return lvl


@pytest.mark.parametrize("kind,n,seed", [('one', 200_000, 31), ('seq', 340_000, 32)])
def test_cf_generators_at_wavefront_depth_4(dev, kind, n, seed):
    """BASELINE config C3 runs 4 dependency levels (1M patients); here the smallest cohorts that reach level 4:
    row counts, applications, sequence lengths (patient ids, current t) bit-exact and volumes 1e-9 against the
    oracle on (a) every row of the global source prefix (levels 1-3, the reference verbatim) and (b) a strided set
    of patients behind it up to the last one (levels 3-4; windowed oracle, pinned to the verbatim one on the CPU).
    Total rows x width exceeds 2^31, so the expanded row ranges sit behind int32 offsets."""
    import torch
    from b200_insite import counterfactual as cf
    from oracle import sim_oracle as so
    T, H = 60, 5
    seq = kind == 'seq'
    W = T + (H if seq else 0)
    params, draws = h.random_cohort(n, seed=seed, extra=H if seq else 0)
    pd_ = dev.to_device(dev.pack_params(params))
    dd = [dev.to_device(draws[k]) for k in ('noise', 'recovery', 'chemo', 'radio')]
    ptypes = dev.to_device(np.asarray(params['patient_types'], dtype=np.float64))
    cohort = cf.sim_cf_treatment_seq(pd_, *dd, T, H) if seq else cf.sim_cf_one_step(pd_, *dd, T)
    torch.cuda.synchronize()
    off = cohort.row_offsets.cpu().numpy()
    assert cohort.levels == 4
    lvl = _cohort_levels(off, n)
    assert lvl.max() == 4
    assert int(off[-1]) == cohort.total_rows and cohort.total_rows * W > 2 ** 31
    exact = CFS_EXACT if seq else CF1_EXACT
    run = (lambda p, d, **kw: so.sim_cf_treatment_seq(p, T, H, d, **kw)) if seq else \
          (lambda p, d, **kw: so.sim_cf_one_step(p, T, d, **kw))
    # (a) the global source prefix, reference semantics verbatim
    P = n // (480 if seq else 200)
    ref = run(_sub_params(params, slice(0, P)), {k: v[:P] for k, v in draws.items()})
    R = ref['cancer_volume'].shape[0]
    assert R >= n and int(off[P]) == R
    got = {k: v.cpu().numpy() for k, v in cf.expand(cohort, ptypes, 0, R).items()}
    _check_cf(got, ref, exact, f"{kind} prefix n={n}")
    del got
    # (b) strided patients behind the prefix, the last ones at level 4
    idx = np.unique(np.concatenate([np.linspace(P, n - 1, 300).astype(np.int64), np.arange(n - 40, n)]))
    assert lvl[idx].max() == 4 and (lvl[idx] == 4).sum() >= 40
    sub = run(_sub_params(params, idx), {k: v[idx] for k, v in draws.items()}, window_rows=ref['cancer_volume'][idx])
    parts = [cf.expand(cohort, ptypes, int(off[i]), int(off[i + 1])) for i in idx]
    got = {k: torch.cat([p[k] for p in parts]).cpu().numpy() for k in parts[0]}
    if seq:   # the windowed oracle numbers the subset's patients 0..len(idx)-1
        sub['patient_ids_all_trajectories'] = idx[sub['patient_ids_all_trajectories'].astype(np.int64)].astype(np.float64)
    assert int(off[idx[-1] + 1]) * W > 2 ** 31
    _check_cf(got, sub, exact, f"{kind} strided n={n}")


@pytest.mark.parametrize("kind,n", [('one_step', 5000), ('treatment_seq', 3001)])
def test_cf_shards_with_source_prefix_equal_the_single_launch(dev, kind, n):
    """SURVEY 8(e) exception (cancer_simulation.py:471, :671): a cohort split into three contiguous shards, each
    simulated in one launch against the redundantly simulated global source prefix (b200i_cf_source, global_base),
    is bit-identical to the single launch with its level loop -- compact arrays and global row offsets."""
    import torch
    from b200_insite import cohort as co, counterfactual as cf
    T, H = 60, 5
    seq = kind == 'treatment_seq'
    params, draws = h.random_cohort(n, seed=44, extra=H if seq else 0)
    pd_ = dev.to_device(dev.pack_params(params))
    dd = [dev.to_device(draws[k]) for k in ('noise', 'recovery', 'chemo', 'radio')]
    whole = cf.sim_cf_treatment_seq(pd_, *dd, T, H) if seq else cf.sim_cf_one_step(pd_, *dd, T)
    torch.cuda.synchronize()
    assert whole.levels >= 3
    P = co.source_prefix_length(whole.row_offsets.cpu().numpy(), n)
    assert 0 < P < n // 3          # the prefix is a small part of the first shard
    world, base, got = 3, 0, []
    for rank in range(world):
        lo, hi = co.shard_bounds(n, rank, world)
        cut = lambda a, s: a[:, s].contiguous() if a.dim() == 2 and a.shape[0] == 10 else a[s].contiguous()
        shard_in = (cut(pd_, slice(lo, hi)),) + tuple(cut(a, slice(lo, hi)) for a in dd)
        prefix_in = (cut(pd_, slice(0, P)),) + tuple(cut(a, slice(0, P)) for a in dd)
        shard, src, _ = cf.sim_cf_shard(kind, T, H, n, lo, shard_in, prefix_in, row_base=base)
        torch.cuda.synchronize()
        assert shard.levels == 1 and src.n == P
        base += shard.total_rows
        got.append(shard)
    assert base == whole.total_rows
    for name in ('factual', 'codes', 'cf', 'valid', 'n_steps', 'n_rows'):
        if getattr(whole, name) is None:
            continue
        cat = torch.cat([getattr(s, name) for s in got])
        ref = getattr(whole, name)
        same = torch.equal(cat.view(torch.uint8) if cat.dtype.is_floating_point else cat,
                           ref.view(torch.uint8) if ref.dtype.is_floating_point else ref)
        assert same, f"{kind}: {name} differs between the shards and the single launch"
    offs = torch.cat([s.row_offsets[:-1] for s in got] + [got[-1].row_offsets[-1:]])
    assert torch.equal(offs, whole.row_offsets)
    # a prefix that does not cover the rows the shard reads is an error, not a silent zero window
    lo, hi = co.shard_bounds(n, 2, world)
    cut = lambda a, s: a[:, s].contiguous() if a.dim() == 2 and a.shape[0] == 10 else a[s].contiguous()
    with pytest.raises((RuntimeError, ValueError)):
        cf.sim_cf_shard(kind, T, H, n, lo, (cut(pd_, slice(lo, hi)),) + tuple(cut(a, slice(lo, hi)) for a in dd),
                        (cut(pd_, slice(0, 2)),) + tuple(cut(a, slice(0, 2)) for a in dd), row_base=0)


def test_sim_factual_side_rejects_a_code_pitch_the_kernel_cannot_store(dev):
    """The side-output kernel writes the code bytes of every 16-column box as one 16-byte word: a pitch of T = 60 (or
    any non-multiple of 16) must be refused, not written past the row."""
    import torch
    params, draws = h.random_cohort(64, seed=3)
    pd_ = dev.to_device(dev.pack_params(params))
    dd = [dev.to_device(draws[k]) for k in ('noise', 'recovery', 'chemo', 'radio')]
    for bad in (60, 62, 48):
        codes = torch.zeros((64, bad), dtype=torch.uint8, device='cuda')
        with pytest.raises(RuntimeError, match="code_pitch"):
            dev.sim_factual_side(pd_, *dd, 60, codes=codes)
    out, codes, pm = dev.sim_factual_side(pd_, *dd, 60)
    torch.cuda.synchronize()
    assert codes.shape[1] == 64
