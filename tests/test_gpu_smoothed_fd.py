"""model.use_smoothed_finite_difference (sindy.py:196-198): the smoothing pre-pass and the fit behind it, against the
oracle's explicit smoothed design matrices (the smoother itself is pinned against scipy in tests/test_oracle.py)."""
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


def _collection(seed, mode):
    from b200_insite.dataset import SyntheticCancerDatasetCollection
    col = SyntheticCancerDatasetCollection(2.0, 2.0, {'train': 1000, 'val': 100, 'test': 100}, seed=seed,
                                           treatment_mode=mode)
    col.process_data_multi()
    return col


def test_smoothing_pass_equals_per_snippet_savgol(dev):
    """Every snippet of every patient, smoothed by the oracle on its own, equals the matching slice of the one smoothed
    row the kernel writes (snippets share only edge samples, which stay as they are)."""
    import torch
    from oracle import sindy_np as sp
    col = _collection(1, 'multiclass')
    tr = col.train_f
    prev = tr.data['prev_outputs'] * tr.scaling_params['output_stds'] + tr.scaling_params['output_means']
    vol = np.concatenate((prev[:, 0].reshape(-1, 1), np.squeeze(tr.data['unscaled_outputs'], -1)), axis=1)
    codes = np.argmax(tr.data['current_treatments'], axis=-1)
    seq = tr.data['sequence_lengths'].astype(np.int64)
    n, T = vol.shape
    chemo = np.zeros((n, T)); radio = np.zeros((n, T))
    chemo[:, :T - 1] = codes & 1; radio[:, :T - 1] = (codes >> 1) & 1
    out = dev.smooth_snippets(dev.to_device(vol), dev.to_device(chemo), dev.to_device(radio),
                              dev.to_device(seq.astype(np.float64)))
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    checked = 0
    for p in range(0, n, 7):
        L, a = int(seq[p]), 0
        for i in range(1, L + 1):
            if i == L or codes[p, i] != codes[p, i - 1]:
                want = sp.savgol_w2_p1(vol[p, a:i + 1])
                np.testing.assert_array_equal(out[p, a:i + 1], want)
                checked += 1
                a = i
        np.testing.assert_array_equal(out[p, L + 1:], vol[p, L + 1:])
    assert checked > 500
    # joint trajectories: columns 1..L of a patient
    outj = dev.smooth_snippets(dev.to_device(vol), dev.to_device(chemo), dev.to_device(radio),
                               dev.to_device(seq.astype(np.float64)), joint=True).cpu().numpy()
    for p in range(0, n, 11):
        L = int(seq[p])
        np.testing.assert_array_equal(outj[p, 1:L + 1], sp.savgol_w2_p1(vol[p, 1:L + 1]))
        assert outj[p, 0] == vol[p, 0]


@pytest.mark.parametrize("joint", [False, True])
def test_fit_with_smoothed_finite_difference_equals_oracle(dev, joint):
    from oracle import sindy_np as sp
    from b200_insite.config import default_config
    from b200_insite.sindy import SINDY
    col = _collection(10 if joint else 1, 'multilabel' if joint else 'multiclass')
    kw = dict(treatment_mode='multilabel', joint_model=True, seed=10) if joint else {}
    model = SINDY(default_config(insite=False, use_smoothed_finite_difference=True, **kw), col)
    model.fit(col.train_f)
    plain = SINDY(default_config(insite=False, **kw), col)
    plain.fit(col.train_f)
    tr = col.train_f
    if joint:
        want, sup, _ = sp.fit_population_joint(tr.data, tr.scaling_params, smoothed=True)
    else:
        want, sup, _ = sp.fit_population(tr.data, tr.scaling_params, smoothed=True)
    assert np.array_equal(model.support_.reshape(-1), np.asarray(sup).reshape(-1))
    np.testing.assert_allclose(model.joint_coefs, want, rtol=1e-7, atol=1e-12)
    assert np.abs(model.joint_coefs - plain.joint_coefs).max() > 1e-4      # the option does change the fit


def test_smoothing_pass_edge_lengths(dev):
    """Sequence length 1 (two samples: nothing to smooth), full length, empty cohort."""
    import torch
    T = 9
    rng = np.random.default_rng(2)
    vol = rng.uniform(1.0, 50.0, size=(3, T))
    chemo = np.zeros((3, T)); radio = np.zeros((3, T))
    seq = np.array([1.0, T - 1.0, 4.0])
    out = dev.smooth_snippets(dev.to_device(vol), dev.to_device(chemo), dev.to_device(radio), dev.to_device(seq)).cpu().numpy()
    np.testing.assert_array_equal(out[0], vol[0])
    want = vol[1].copy(); want[1:T - 1] = 0.5 * vol[1, 1:T - 1] + 0.5 * vol[1, 2:]
    np.testing.assert_array_equal(out[1], want)
    want = vol[2].copy(); want[1:4] = 0.5 * vol[2, 1:4] + 0.5 * vol[2, 2:5]
    np.testing.assert_array_equal(out[2], want)
    z = torch.zeros((0, T), dtype=torch.float64, device='cuda')
    assert dev.smooth_snippets(z, z, z, torch.zeros(0, dtype=torch.float64, device='cuda')).shape == (0, T)
