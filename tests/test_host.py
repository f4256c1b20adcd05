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
