"""CPU tests of the host layer: parameter generation, scaling parameters, C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers as h


def test_generate_params_bit_exact_and_same_rng_consumption(golden_dir):
    import b200_insite.cancer_simulation as cs
    g = h.load_npz('ref_sim_small.npz')
    np.random.seed(7)
    p = cs.generate_params(192, 2.0, 2.0, 15, 0)
    for k in h.PARAM_KEYS:
        assert np.array_equal(p[k], g[f'train/params/{k}']), k
        assert p[k].dtype == g[f'train/params/{k}'].dtype, k
    assert np.array_equal(p['initial_stages'], g['train/params/initial_stages'])
    assert p['window_size'] == 15 and p['lag'] == 0
    # the stream continues exactly where the reference left it: the next subset's parameters match too
    np.random.rand(192, 60 * 4 // 4); np.random.seed(7)
    inputs = h.collection_inputs(7, 2.0, 192, 24, 24)
    np.random.seed(7)
    p1 = cs.generate_params(192, 2.0, 2.0, 15, 0)
    from oracle import rng_export as rx
    rx.draw_factual(192, 60)
    p2 = cs.generate_params(24, 2.0, 2.0, 15, 0)
    for k in h.PARAM_KEYS:
        assert np.array_equal(p2[k], inputs['val'][0][k]), k


def test_get_scaling_params_matches_reference_fixture():
    import b200_insite.cancer_simulation as cs
    g = h.load_npz('ref_sim_small.npz')
    sim = {k: g[f'train/out/{k}'] for k in ('cancer_volume', 'chemo_dosage', 'radio_dosage', 'sequence_lengths',
                                            'patient_types')}
    means, stds = cs.get_scaling_params(sim)
    assert list(means.index) == ['cancer_volume', 'chemo_dosage', 'radio_dosage', 'patient_types']
    assert list(means.values) == list(g['train/scaling_means'])
    assert list(stds.values) == list(g['train/scaling_stds'])


def test_constants_match_reference_values():
    import b200_insite.cancer_simulation as cs
    from b200_insite import device as dev
    assert cs.TUMOUR_DEATH_THRESHOLD == 1150.3465099894624
    assert cs.calc_volume(30) == 14137.166941154068
    assert cs.calc_diameter(cs.TUMOUR_DEATH_THRESHOLD) == 12.999999999999998
    c = dev.sim_consts(15, 0)
    assert c.drug_decay == 0.5 and c.sphere_coef == 4.1887902047863905 and c.death_threshold == 1150.3465099894624
    assert dev.STANDARD_DT == 0.16666666666666666


def _declared_symbols():
    hdr = open(os.path.join(h.ROOT, 'include', 'b200i.h')).read()
    return sorted(set(re.findall(r'B200I_API[^;(]*?\b(b200i_\w+)\s*\(', hdr)))


def test_c_abi_library_exports_every_declared_symbol():
    """The shared library loads (no GPU needed) and exports exactly what include/b200i.h declares."""
    import b200_insite._native as nat
    if not os.path.isfile(nat.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    names = _declared_symbols()
    assert len(names) >= 15
    lib = ctypes.CDLL(nat.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200i.h but not exported"
    assert sorted(nat.SIGNATURES) == names, "python binding table and header disagree"
    lib.b200i_version.restype = ctypes.c_int
    assert lib.b200i_version() >= 100


def test_argument_errors_are_reported_without_a_gpu():
    import b200_insite._native as nat
    lib = nat.load()
    rc = lib.b200i_stlsq_population(None, 1e-3, 0.5, 100, None, None, None)
    assert rc == -1
    assert b"NULL" in lib.b200i_last_error()
    with pytest.raises(RuntimeError, match="status -1"):
        nat.check(rc, "b200i_stlsq_population")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(h.ROOT, 'ode-discovery-for-longitudinal-heterogeneous-treatment-effects-inference_b200')
    for root, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(root, fn)).read()
                assert 'oracle' not in src.replace('# oracle', ''), f"{fn} mentions the oracle"


@pytest.mark.parametrize("mode", ["multiclass", "multilabel"])
def test_dataset_transforms_equal_the_restatement(mode):
    """process_data / process_sequential_test / process_sequential_multi of the product's dataset
    classes (vectorised numpy) vs the loop-for-loop restatement of dataset.py:92-192, 395-473, 533-552, on the
    oracle's simulator output (no GPU needed: the dataset objects are filled in by hand)."""
    import warnings
    from oracle import sindy_np as sp
    from b200_insite.dataset import SyntheticCancerDataset, SyntheticCancerDatasetCollection
    from b200_insite.cancer_simulation import TUMOUR_DEATH_THRESHOLD
    warnings.filterwarnings('ignore')
    inputs = h.collection_inputs(5, 2.0, 96, 16, 16)
    o = h.oracle_collection(inputs)

    def mk(data, name):
        d = object.__new__(SyntheticCancerDataset)
        d.data = dict(data); d.subset_name = name
        d.processed = d.processed_sequential = d.processed_autoregressive = False
        d.treatment_mode = mode; d.exploded = False; d.norm_const = TUMOUR_DEATH_THRESHOLD
        return d
    col = object.__new__(SyntheticCancerDatasetCollection)
    col.train_f, col.val_f = mk(o['train'], 'train'), mk(o['val'], 'val')
    col.test_cf_one_step, col.test_cf_treatment_seq = mk(o['one'], 'test'), mk(o['seq'], 'test')
    col.projection_horizon = 5
    col.train_scaling_params = col.train_f.get_scaling_params()
    means, stds = col.train_f.get_scaling_params()
    col.process_data_multi()
    keys = ('prev_treatments', 'current_treatments', 'current_covariates', 'outputs', 'active_entries',
            'unscaled_outputs', 'prev_outputs', 'static_features')
    for ds, name in ((col.train_f, 'train'), (col.val_f, 'val'), (col.test_cf_one_step, 'one')):
        want, sc = sp.process_data(o[name], means, stds, treatment_mode=mode)
        for k in keys:
            assert ds.data[k].shape == want[k].shape and np.array_equal(ds.data[k], want[k]), (name, k)
        for k in sc:
            assert np.array_equal(np.asarray(ds.scaling_params[k]), np.asarray(sc[k])), k
    want, sc = sp.process_data(o['seq'], means, stds, treatment_mode=mode)
    seq = col.test_cf_treatment_seq
    for k in keys:                                   # after process_sequential_multi .data is the full-length form again
        assert np.array_equal(seq.data[k], want[k]), ('seq', k)
    assert np.array_equal(seq.data['future_past_split'], o['seq']['sequence_lengths'] - 5)
    want5 = sp.process_sequential_test(want, sc, 5)
    for k in want5:
        assert np.array_equal(seq.data_processed_seq[k], want5[k]), ('seq5', k)
    assert 'future_past_split' not in seq.data_original
    # the reference caches collections with shelve (run_utils.py:4-19): a pickled + restored collection is equal, and a
    # restored dataset that is processed only afterwards still works (row builders are closures and stay behind)
    import pickle
    back = pickle.loads(pickle.dumps(col))
    for k in keys:
        assert np.array_equal(back.test_cf_treatment_seq.data[k], want[k]), ('pickled', k)
    assert np.array_equal(back.test_cf_treatment_seq.data_processed_seq['outputs'], want5['outputs'])
    fresh = mk(o['seq'], 'test')
    fresh.process_data(col.train_f.get_scaling_params())
    fresh = pickle.loads(pickle.dumps(fresh))
    fresh.process_sequential_test(5)
    for k in want5:
        assert np.array_equal(fresh.data[k], want5[k]), ('restored', k)


def test_lazy_dataset_dictionary_behaves_like_a_dict():
    """dataset.LazyDict: pending arrays are ordinary keys for indexing / `in` / get, enumeration and pickling build
    them, copies stay lazy and share arrays, assigning a watched key fires its hook."""
    import copy
    import pickle
    from b200_insite.dataset import LazyDict
    calls = []
    d = LazyDict(a=np.arange(3))
    d.set_lazy('b', lambda: calls.append('b') or np.ones(2))
    fired = []
    d.on_set('b', lambda: fired.append(1))
    assert 'b' in d and len(d) == 2 and not calls
    c = d.copy()
    assert isinstance(c, LazyDict) and c['a'] is d['a'] and not calls
    assert d.get('zz', 7) == 7 and d.get('b').sum() == 2 and calls == ['b']
    assert d['b'] is d['b'] and calls == ['b']                       # built once
    assert sorted(c.keys()) == ['a', 'b'] and calls == ['b', 'b']    # the copy builds its own on enumeration
    p = pickle.loads(pickle.dumps(c))
    assert type(p) is dict and sorted(p) == ['a', 'b']
    e = LazyDict(x=1)
    e.set_lazy('y', lambda: [1, 2])
    assert type(copy.deepcopy(e)) is dict and dict(e) == {'x': 1, 'y': [1, 2]}
    d['b'] = np.zeros(1)
    assert fired == [1] and d['b'].shape == (1,)
    with pytest.raises(KeyError):
        d['missing']
    del c['b']
    assert 'b' not in c


def test_continuous_generate_params_bit_exact_and_rng_consumption():
    """EQ_5 parameter generator (continuous/continuous.py:68-224) vs the committed vectors of the unmodified reference
    (oracle/make_golden_continuous.py) and, where /root/reference exists, vs the live reference for all four equations:
    every array bit-exact and the global RNG left in the same state."""
    import hashlib
    from b200_insite import continuous as ct
    g = h.load_npz('ref_continuous_small.npz')
    for eq in ('EQ_5_A', 'EQ_5_D'):
        np.random.seed(17)
        p = ct.generate_params(40, 2.0, 2.0, 15, 0, ct.Equation[eq])
        for k, v in p.items():
            ref = g[f'{eq}/factual/params/{k}']
            assert np.array_equal(np.asarray(v), ref), (eq, k)
        assert set(p) == {k.split('/')[-1] for k in g.files if k.startswith(f'{eq}/factual/params/')}
    assert p['observation_noise'] == 0.01 and len(np.unique(p['beta_c'])) > 3
    from oracle import ref_loader
    if not ref_loader.reference_available():
        return
    import sys
    ref = ref_loader.load_reference_continuous()
    Eq = sys.modules["src.data.pkpd.pkpd_simulation"].Equation
    for eq in ('EQ_5_A', 'EQ_5_B', 'EQ_5_C', 'EQ_5_D'):
        np.random.seed(3)
        a = ref.generate_params(257, 3.0, 1.0, 15, 0, Eq[eq])
        sa = hashlib.sha256(np.random.get_state()[1].tobytes()).hexdigest()
        np.random.seed(3)
        b = ct.generate_params(257, 3.0, 1.0, 15, 0, ct.Equation[eq])
        sb = hashlib.sha256(np.random.get_state()[1].tobytes()).hexdigest()
        assert sa == sb and set(a) == set(b)
        for k in a:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (eq, k)
    with pytest.raises(ValueError):
        ct.generate_params(4, 2.0, 2.0, 15, 0, ct.Equation.EQ_4_A)


def test_model_switches_outside_the_built_path_are_refused_up_front():
    """The mirror of SINDY raises NotImplementedError in its constructor (no GPU needed) for the reference switches that
    are not built, and accepts the ones that are (sindy.py:95-130 of the package)."""
    import pytest
    from b200_insite.config import default_config
    from b200_insite.sindy import SINDY
    ok = [dict(insite=False), dict(insite=True), dict(insite=False, use_smoothed_finite_difference=True),
          dict(insite=True, use_smoothed_finite_difference=True), dict(insite=False, sindy_quantize=True),
          dict(insite=False, ablation_more_complex_basis_functions=True),
          dict(insite=True, joint_model=True, treatment_mode='multilabel')]
    for kw in ok:
        SINDY(default_config(**kw), None, autoregressive=True, has_vitals=False)
    refused = [dict(wsindy=True), dict(smooth_input_data=True),
               dict(insite=True, ablation_more_complex_basis_functions=True),
               dict(insite=False, ablation_more_complex_basis_functions=True, joint_model=True, treatment_mode='multilabel'),
               dict(insite=False, joint_model=True),                          # the joint model is multilabel
               dict(insite=False, treatment_mode='multilabel')]
    for kw in refused:
        with pytest.raises(NotImplementedError):
            SINDY(default_config(**kw), None, autoregressive=True, has_vitals=False)
