"""GPU tests of the drop-in classes on device-resident counterfactual test sets (BASELINE config C1 without the dense
round trip): the dictionaries of simulate_counterfactual_* keep their dense (R, width) arrays pending, every array
process_data / process_sequential_test derive from them stays pending too, and SINDY's two RMSE methods evaluate the
compact cohort.  Anything that reads a dense key gets the reference's arrays (bit-identical to the eager path); anything
that reassigns one drops the compact cohort and the dense path runs."""
import numpy as np
import pytest

import helpers as h

pytestmark = pytest.mark.gpu
RMSE_KEYS = ('encoder_test_rmse_all', 'encoder_test_rmse_orig', 'encoder_test_rmse_last') + \
    tuple(f'decoder_test_rmse_{k}-step' for k in range(2, 7))


@pytest.fixture(scope="module")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from b200_insite import device
    device.require_cuda()
    return device


def _collection(seed=1, sizes=(1000, 100, 100), mode='multiclass'):
    from b200_insite.dataset import SyntheticCancerDatasetCollection
    col = SyntheticCancerDatasetCollection(2.0, 2.0, {'train': sizes[0], 'val': sizes[1], 'test': sizes[2]}, seed=seed,
                                           treatment_mode=mode)
    col.process_data_multi()
    return col


def test_lazy_dictionaries_hold_the_reference_arrays(dev):
    import b200_insite.cancer_simulation as cs
    for kind in ('one', 'seq'):
        np.random.seed(5)
        p = cs.generate_params(37, 2.0, 2.0, 15, 0)
        state = np.random.get_state()
        eager = cs.simulate_counterfactual_1_step(p, 60) if kind == 'one' else cs.simulate_counterfactuals_treatment_seq(p, 60, 5)
        np.random.set_state(state)
        lazy = cs.simulate_counterfactual_1_step(p, 60, lazy=True) if kind == 'one' else \
            cs.simulate_counterfactuals_treatment_seq(p, 60, 5, lazy=True)
        assert lazy.pending('cancer_volume') and lazy.attrs['compact'][0].total_rows == eager['cancer_volume'].shape[0]
        for k in ('sequence_lengths', 'patient_types') + (() if kind == 'one' else ('patient_ids_all_trajectories', 'patient_current_t')):
            assert not lazy.pending(k) and lazy[k].dtype == eager[k].dtype and np.array_equal(lazy[k], eager[k]), k
        assert lazy.pending('chemo_application')
        assert set(lazy.keys()) == set(eager.keys())            # enumerating builds the dense arrays (one expansion)
        for k in eager:
            assert np.array_equal(lazy[k], eager[k]), k
        assert not lazy.pending('cancer_volume')


def test_class_metrics_from_the_compact_cohorts_equal_the_dense_path(dev):
    from b200_insite.config import default_config
    from b200_insite.sindy import run_experiment
    log = h.load_json('ref_log_seed1.json')['sindy']
    col = _collection()
    one, seq = col.test_cf_one_step, col.test_cf_treatment_seq
    assert one.compact_ is not None and seq.compact_ is not None
    assert len(one) == 22748 and len(seq) == 56239
    res, model = run_experiment(default_config(insite=False), col)
    # nothing on the way read a dense row
    assert one.data.pending('cancer_volume') and one.data.pending('prev_outputs') and seq.data.pending('outputs')
    assert seq.data_processed_seq.pending('unscaled_outputs')
    ref = [log['encoder_test_rmse_all'], log['encoder_test_rmse_orig'], log['encoder_test_rmse_last']] + \
        list(log['decoder_test_rmse_2_to_6_step'])
    np.testing.assert_allclose([res[k] for k in RMSE_KEYS], ref, rtol=1e-9)
    # the dense path on the same collection (materialises the rows)
    res_d, _ = run_experiment(default_config(insite=False, compact_evaluation=False), col)
    assert not one.data.pending('cancer_volume') and not seq.data_processed_seq.pending('unscaled_outputs')
    np.testing.assert_allclose([res_d[k] for k in RMSE_KEYS], [res[k] for k in RMSE_KEYS], rtol=1e-10)
    # reassigning a key of the processed dictionary makes the compact cohort stale: it is dropped
    one.data['outputs'] = one.data['outputs'].copy()
    assert one.compact_ is None and seq.compact_ is not None
    res_t, _ = run_experiment(default_config(insite=False), col)
    np.testing.assert_allclose([res_t[k] for k in RMSE_KEYS], [res[k] for k in RMSE_KEYS], rtol=1e-10)


def test_insite_class_metrics_compact_vs_dense(dev):
    from b200_insite.config import default_config
    from b200_insite.sindy import run_experiment
    ins = h.load_json('ref_log_seed1.json')['insite']
    col = _collection()
    res, model = run_experiment(default_config(insite=True, insite_zoom_failure_fallback=False), col)
    assert col.test_cf_one_step.data.pending('cancer_volume')
    assert model.last_fit_info['fits'] < 6000 and model.last_fit_info['rows_covered'] == 56239
    ref = [ins['encoder_test_rmse_all'], ins['encoder_test_rmse_orig'], ins['encoder_test_rmse_last']] + \
        list(ins['decoder_test_rmse_2_to_6_step'])
    np.testing.assert_allclose([res[k] for k in RMSE_KEYS], ref, rtol=1e-9)
    res_d, _ = run_experiment(default_config(insite=True, insite_zoom_failure_fallback=False, compact_evaluation=False), col)
    # same estimator on the same fit problems; the dense rows carry the scaling round trip (x - mu) / sigma * sigma + mu
    np.testing.assert_allclose([res_d[k] for k in RMSE_KEYS], [res[k] for k in RMSE_KEYS], rtol=1e-6)
    for fb in (True, False):      # the fallback of the reference's current code, both paths
        a, _ = run_experiment(default_config(insite=True, insite_zoom_failure_fallback=fb), col)
        b, _ = run_experiment(default_config(insite=True, insite_zoom_failure_fallback=fb, compact_evaluation=False), col)
        np.testing.assert_allclose([a[k] for k in RMSE_KEYS], [b[k] for k in RMSE_KEYS], rtol=1e-6)
    # north-star estimator (ridge-to-prior STLSQ, one fit per (patient, t) by running sums)
    cfg = default_config(insite=True, individualisation='ridge_prior_stlsq')
    r1, _ = run_experiment(cfg, col)
    r2, _ = run_experiment(default_config(insite=True, individualisation='ridge_prior_stlsq', compact_evaluation=False), col)
    np.testing.assert_allclose([r1[k] for k in RMSE_KEYS], [r2[k] for k in RMSE_KEYS], rtol=1e-8)


def test_joint_model_population_metrics_from_compact_cohorts(dev):
    from b200_insite.config import default_config
    from b200_insite.sindy import run_experiment
    log = h.load_json('ref_log_joint_seed10.json')
    col = _collection(seed=log['seed'], mode='multilabel')
    res, model = run_experiment(default_config(insite=False, joint_model=True, treatment_mode='multilabel'), col)
    assert col.test_cf_one_step.data.pending('cancer_volume')
    res_d, _ = run_experiment(default_config(insite=False, joint_model=True, treatment_mode='multilabel',
                                             compact_evaluation=False), col)
    np.testing.assert_allclose([res[k] for k in RMSE_KEYS], [res_d[k] for k in RMSE_KEYS], rtol=1e-10)


def _parse_equation(s):
    """'Treatment k: x_dot = +c*1+c*x0+...' | ... -> list of [(coefficient, term name)] per treatment."""
    import re
    out = []
    for part in s.split(' | '):
        head, body = part.split(' = ')
        assert re.fullmatch(r'Treatment \d: x_dot', head), head
        out.append([(float(c), name) for c, name in re.findall(r'\+(-?[0-9.]+(?:e-?[0-9]+)?)\*([a-z0-9]+(?:\*[a-z0-9]+)*)', body)])
        assert ''.join(f'+{repr(c)}*{name}' for c, name in out[-1]) == body     # nothing but '+repr(float)*term' pieces
    return out


def test_result_log_line_is_the_reference_wire_format(dev):
    """run.py:119-121 / train_sindy.py:72-112: the '[Exp evaluation complete] {...}' line of the seed-1 SINDy and INSITE
    runs parses (with the reference's own parser logic, utils/results_utils.py:126-131) to the dictionary of the
    reference's committed log: same keys in the same order, RMSEs 1e-8 / 1e-9, and the WHOLE global_equation_string:
    same terms in the same order, coefficients printed with repr(float), values 1e-9."""
    import ast
    from b200_insite import runner
    gold = h.load_json('ref_logline_seed1.json')
    for method, overrides, tol in (('sindy', {}, 1e-8), ('insite', dict(insite_zoom_failure_fallback=False), 1e-9)):
        res = runner.run_exp_wrapper_outer(('cancer_sim', method, 1, 2.0), **overrides)
        line = runner.result_log_line(res)
        assert line.startswith('[Exp evaluation complete] {')
        got = ast.literal_eval(line.split('[Exp evaluation complete] ')[1].strip())
        ref = ast.literal_eval(gold[method])
        assert list(got.keys()) == list(ref.keys())
        for k, v in ref.items():
            if k == 'seconds_taken':
                assert got[k] < v                       # 13.1 s / 84.8 s in the reference's log
            elif k == 'global_equation_string':
                g, r = _parse_equation(got[k]), _parse_equation(v)
                assert [[n for _, n in t] for t in g] == [[n for _, n in t] for t in r]
                np.testing.assert_allclose([c for t in g for c, _ in t], [c for t in r for c, _ in t], rtol=1e-9)
            elif isinstance(v, float):
                np.testing.assert_allclose(got[k], v, rtol=tol, err_msg=k)
            else:
                assert got[k] == v, k
    bad = runner.run_exp_wrapper_outer(('EQ_4_A', 'sindy', 1, 2.0))
    assert bad['errored'] is True and bad['dataset_name'] == 'EQ_4_A'


def test_dataset_cache_round_trip(dev, tmp_path):
    """run_utils.get_dataset: shelve cache keyed by str(args.dataset); a cached collection gives the same results."""
    from b200_insite import runner
    from b200_insite.config import default_config
    args = default_config(insite=False, n_train=300, n_val=30, n_test=30, seed=4)
    args.force_recache = True
    path = str(tmp_path / "ct_datasets")
    col = runner.get_dataset(args, cache_path=path)
    r1 = runner.main(args, col)
    args.force_recache, args.load_from_cache = False, True
    col2 = runner.get_dataset(args, cache_path=path)
    assert col2 is not col
    r2 = runner.main(args, col2)
    for k in r1:
        if isinstance(r1[k], float):
            np.testing.assert_allclose(r2[k], r1[k], rtol=1e-10)
    assert r1['global_equation_string'] == r2['global_equation_string']


def test_quantised_population_model(dev):
    """model.sindy_quantize (pkpd/utils.py:389-390): the integrated expression and the logged string carry the
    coefficients rounded to sindy_quantize_global_model_round_to decimals (terms are still selected by |c| > 1e-3 of
    the un-rounded value)."""
    from b200_insite.config import default_config
    from b200_insite.sindy import run_experiment
    from oracle import sindy_np as sp
    col = _collection()
    base, m0 = run_experiment(default_config(insite=False), col)
    res, model = run_experiment(default_config(insite=False, sindy_quantize=True, sindy_quantize_global_model_round_to=2), col)
    np.testing.assert_array_equal(model.joint_coefs, m0.joint_coefs)            # the fit itself is not quantised
    for coef, name in [t for tr in _parse_equation(res['global_equation_string']) for t in tr]:
        assert coef == round(coef, 2)
    # dense restatement with the rounded coefficients
    ds = col.test_cf_one_step
    scp = ds.scaling_params
    prev = np.squeeze(ds.data['prev_outputs'] * scp['output_stds'] + scp['output_means'], -1)
    static = (ds.data['static_features'] * scp['inputs_stds'][1:2] + scp['input_means'][1:2])[:, 0]
    codes = np.argmax(ds.data['current_treatments'], -1)
    c = m0.joint_coefs
    ce = np.where(np.abs(c) > 1e-3, np.round(c, 2), 0.0)
    un = sp.rollout_unscaled(prev[:, 0], codes, static, ce)
    scaled = ((un - scp['output_means']) / scp['output_stds'])[..., None]
    orig, all_, last = sp.masked_rmse(scaled, ds.data, scp)
    np.testing.assert_allclose([res['encoder_test_rmse_orig'], res['encoder_test_rmse_all'], res['encoder_test_rmse_last']],
                               [orig, all_, last], rtol=1e-9)
    dense, _ = run_experiment(default_config(insite=False, sindy_quantize=True, sindy_quantize_global_model_round_to=2,
                                             compact_evaluation=False), col)
    np.testing.assert_allclose([dense[k] for k in RMSE_KEYS], [res[k] for k in RMSE_KEYS], rtol=1e-10)
    assert abs(res['encoder_test_rmse_all'] - base['encoder_test_rmse_all']) > 1e-6


def test_second_population_known_answer_seed10(dev):
    """results/2_main_table/final_with_insite.txt:2094: population SINDy on the collection seeded with 10 (the run's
    dataset cache held that collection): the 8 RMSEs and the 16 coefficients of the logged equation string."""
    import ast
    from b200_insite import runner
    ref = ast.literal_eval(h.load_json('ref_logline_seed1.json')['sindy_seed10'])
    col = _collection(seed=10)
    res = runner.main(__import__('b200_insite.config', fromlist=['default_config']).default_config(insite=False, seed=10), col)
    for k in RMSE_KEYS:
        np.testing.assert_allclose(res[k], ref[k], rtol=1e-8, err_msg=k)
    g, r = _parse_equation(res['global_equation_string']), _parse_equation(ref['global_equation_string'])
    assert [[n for _, n in t] for t in g] == [[n for _, n in t] for t in r]
    np.testing.assert_allclose([c for t in g for c, _ in t], [c for t in r for c, _ in t], rtol=1e-8)
