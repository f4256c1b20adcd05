"""Experiment wrapper of the reference without hydra / MLflow: the pieces of run.py and libs_m/ct/runnables that sit
between the configuration and the model.

    get_dataset(args)                      runnables/run_utils.py:4-19   (shelve cache keyed by str(args.dataset))
    main(args)                             runnables/train_sindy.py:27-113 (dataset -> process_data_multi -> SINDY.fit ->
                                           one-step / n-step RMSEs -> result dict, same keys in the same order)
    run_exp_ct(...)                        run.py:174-306 for the sindy / insite backbones (config overrides -> main ->
                                           'method', 'seed', 'seconds_taken')
    run_exp_wrapper_outer(args, config)    run.py:154-170 ('errored' flag, 'dataset_name', 'seed', 'method_name',
                                           'domain_conf'; exceptions become {'errored': True})
    result_log_line(result)                run.py:119-121: '[Exp evaluation complete] {...}', the line
                                           utils/results_utils.py:121-128 parses back with ast.literal_eval
"""
import logging
import shelve
import time
import traceback

import numpy as np

from .config import default_config

logger = logging.getLogger(__name__)
METHODS = {'sindy': dict(insite=False), 'insite': dict(insite=True)}      # config/backbone/{sindy,insite}.yaml


def get_dataset(args, cache_path="ct_datasets"):
    """run_utils.py:4-19.  The cache key ignores nothing the reference does not ignore: str(args.dataset)."""
    from .dataset import SyntheticCancerDatasetCollection
    d = args.dataset

    def build():
        return SyntheticCancerDatasetCollection(
            d.chemo_coeff, d.radio_coeff, {'train': d.num_patients.train, 'val': d.num_patients.val,
                                           'test': d.num_patients.test}, seed=d.seed, window_size=d.window_size,
            max_seq_length=d.max_seq_length, projection_horizon=d.projection_horizon, lag=d.lag,
            cf_seq_mode=d.cf_seq_mode, treatment_mode=d.treatment_mode)
    record = str(args.dataset)
    if args.get('force_recache', False):
        col = build()
        with shelve.open(cache_path) as db:
            db[record] = col
        return col
    if args.get('load_from_cache', False):
        try:
            with shelve.open(cache_path) as db:
                return db[record]
        except KeyError:
            col = build()
            with shelve.open(cache_path) as db:
                db[record] = col
            return col
    return build()


def main(args, dataset_collection=None):
    """train_sindy.main (:27-113) -> the result dictionary."""
    from .sindy import SINDY
    results = {}
    np.random.seed(args.exp.seed)                       # seed_everything (:37); the collection re-seeds with dataset.seed
    col = get_dataset(args) if dataset_collection is None else dataset_collection
    col.process_data_multi()
    model = SINDY(args, col)
    model.fit(col.train_f, col.val_f)
    if hasattr(col, 'test_cf_one_step'):
        orig, all_, last = model.get_normalised_masked_rmse(col.test_cf_one_step, one_step_counterfactual=True)
        logger.info(f'Test normalised RMSE (all): {all_}; Test normalised RMSE (orig): {orig}; '
                    f'Test normalised RMSE (only counterfactual): {last}')
        results.update({'encoder_test_rmse_all': all_, 'encoder_test_rmse_orig': orig, 'encoder_test_rmse_last': last})
    test_rmses = {}
    if hasattr(col, 'test_cf_treatment_seq'):
        test_rmses = model.get_normalised_n_step_rmses(col.test_cf_treatment_seq)
    test_rmses = {f'{k + 2}-step': v for (k, v) in enumerate(test_rmses)}
    logger.info(f'Test normalised RMSE (n-step prediction): {test_rmses}')
    results.update({('decoder_test_rmse_' + k): v for (k, v) in test_rmses.items()})
    results.update({'global_equation_string': model.global_equation_string, 'fine_tuned': model.insite})
    return results


def run_exp_ct(dataset_name, method_name, seed, domain_conf, config=None, dataset_collection=None, **model_overrides):
    """run.py:174-306 for dataset_name 'cancer_sim' and the sindy / insite backbones.  `config` carries the outer keys
    the reference reads (run.train_samples / val_samples / test_samples, sindy.sindy_alpha, the per-dataset
    sindy_threshold and lam); defaults = config/config.yaml."""
    if dataset_name != 'cancer_sim' or method_name not in METHODS:
        raise NotImplementedError(f"{dataset_name}/{method_name}: the accelerated path covers cancer_sim with sindy / insite")
    c = dict(train_samples=1000, val_samples=100, test_samples=100, sindy_alpha=0.5, sindy_threshold=1e-3, lam=10.0)
    c.update(config or {})
    t00 = time.perf_counter()
    domain_conf = int(domain_conf)                      # run.py:183
    args = default_config(gamma=float(domain_conf), seed=seed, n_train=c['train_samples'], n_val=c['val_samples'],
                          n_test=c['test_samples'], sindy_alpha=c['sindy_alpha'], sindy_threshold=c['sindy_threshold'],
                          lam=c['lam'], **METHODS[method_name], **model_overrides)
    result = main(args, dataset_collection)
    result.update({'method': method_name, 'seed': seed, 'seconds_taken': time.perf_counter() - t00})
    return result


def run_exp_wrapper_outer(args, config=None, debug_mode=False, **kwargs):
    """run.py:154-170."""
    dataset_name, method_name, seed, domain_conf = args
    logger.info(f'[Now evaluating exp] {args}')
    try:
        result = run_exp_ct(dataset_name, method_name, seed, domain_conf, config=config, **kwargs)
        result['errored'] = False
    except Exception as e:      # noqa: BLE001  (the reference logs and carries on, :162-168)
        if debug_mode:
            raise
        logger.exception(f'[Error] {e}')
        logger.info(f"[Failed evaluating exp] {args}\t| error={e}")
        traceback.print_exc()
        result = {'errored': True}
    result.update({'dataset_name': dataset_name, 'seed': seed, 'method_name': method_name, 'domain_conf': domain_conf})
    return result


def result_log_line(result):
    """run.py:119-121."""
    # numpy >= 2 prints scalars as np.float64(...): the reference's parser (ast.literal_eval) needs plain numbers
    printable = {k: v.tolist() if isinstance(v, (np.ndarray, np.generic)) else v for k, v in result.items()}
    return f'[Exp evaluation complete] {printable}'
