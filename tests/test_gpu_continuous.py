"""GPU tests of the EQ_5 simulator drop-ins (b200_insite.continuous <- libs_m/ct/src/data/continuous/continuous.py)
against vectors of the unmodified reference (tests/golden/ref_continuous_small.npz, oracle/make_golden_continuous.py):
applications, dose channel, sequence lengths, patient ids / current t bit-exact, volumes 1e-9 (the observation noise is
drawn from the global numpy stream in the reference's order, so it is the same noise), RNG state identical afterwards."""
import hashlib

import numpy as np
import pytest

import helpers as h

pytestmark = pytest.mark.gpu
T, H = 30, 5


@pytest.fixture(scope="module")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from b200_insite import device
    device.require_cuda()
    return device


@pytest.mark.parametrize("eq", ['EQ_5_A', 'EQ_5_D'])
def test_continuous_simulators_match_reference_vectors(dev, eq):
    from b200_insite import continuous as ct
    g = h.load_npz('ref_continuous_small.npz')
    E = ct.Equation[eq]
    np.random.seed(17)
    for kind, n in (('factual', 40), ('one', 6), ('seq', 5)):
        p = ct.generate_params(n, 2.0, 2.0, 15, 0, E)
        if kind == 'factual':
            sim = ct.simulate_factual(p, T, E)
        elif kind == 'one':
            sim = ct.simulate_counterfactual_1_step(p, T, E)
        else:
            sim = ct.simulate_counterfactuals_treatment_seq(p, T, H, E)
        ref = {k.split('/')[-1]: g[k] for k in g.files if k.startswith(f'{eq}/{kind}/out/')}
        assert list(sim.keys()) == list(ref.keys()) or set(sim) == set(ref)
        for k, r in ref.items():
            got = np.asarray(sim[k])
            assert got.shape == r.shape, (eq, kind, k, got.shape, r.shape)
            if k in ('cancer_volume', 'chemo_probabilities', 'radio_probabilities'):
                np.testing.assert_allclose(got, r, rtol=1e-9, atol=1e-12, err_msg=f'{eq} {kind} {k}')
            else:
                assert np.array_equal(got, r), (eq, kind, k)
        state = hashlib.sha256(np.random.get_state()[1].tobytes()).hexdigest()
        assert state == str(g[f'{eq}/{kind}/rng_after']), (eq, kind, 'global RNG state')
