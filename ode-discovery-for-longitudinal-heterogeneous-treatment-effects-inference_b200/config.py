"""Minimal attribute-access configuration tree carrying the ~20 keys of the reference's hydra config that
the SINDy/INSITE path reads: config/config.yaml, config/ct_config.yaml,
config/backbone/insite.yaml, config/dataset/cancer_sim.yaml.  An omegaconf DictConfig composed by the
reference's run.py works in its place (only attribute access is used)."""
from types import SimpleNamespace


class Config(SimpleNamespace):
    def __getitem__(self, k):
        return getattr(self, k)

    def __contains__(self, k):
        return hasattr(self, k)

    def get(self, k, default=None):
        return getattr(self, k, default)


def _tree(d):
    return Config(**{k: _tree(v) if isinstance(v, dict) else v for k, v in d.items()})


def default_config(insite=True, gamma=2.0, seed=1, n_train=1000, n_val=100, n_test=100, treatment_mode='multiclass',
                   **model_overrides):
    """Defaults of the reference: sindy_alpha 0.5 (config.yaml:18), sindy_threshold 1e-3 (:21), lam 10 (:25),
    window 15, lag 0, T 60, H 5, sliding_treatment (cancer_sim.yaml:13-17), insite.yaml:3-23."""
    model = dict(name='INSITE' if insite else 'SINDy', lag_features=1, insite_val_error_threshold=1e-4, lam=10.0,
                 sindy_threshold=1e-3, sindy_alpha=0.5, smooth_input_data=False, sindy_quantize=False,
                 sindy_quantize_global_model_round_to=2, joint_model=False, insite=insite, wsindy=False,
                 use_smoothed_finite_difference=False, tune_hparams=False, ablation_more_complex_basis_functions=False,
                 insight_recover_parametric_dist=False, normalize=False, dataset_name='cancer_sim',
                 dim_treatments=4, dim_vitals=0, dim_static_features=1, dim_outcomes=1,
                 # b200_insite extension: which individualisation estimator INSITE uses
                 individualisation='bfgs_rollout', ridge_prior_lam=1e4)
    model.update(model_overrides)
    return _tree({
        'model': model,
        'dataset': dict(name='tumor_generator', coeff=gamma, chemo_coeff=gamma, radio_coeff=gamma, seed=seed,
                        num_patients=dict(train=n_train, val=n_val, test=n_test), window_size=15, lag=0,
                        max_seq_length=60, projection_horizon=5, cf_seq_mode='sliding_treatment',
                        val_batch_size=512, treatment_mode=treatment_mode),
        'exp': dict(seed=seed, unscale_rmse=True, percentage_rmse=True, logging=False),
        'force_recache': False, 'load_from_cache': False})
