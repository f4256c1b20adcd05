"""Host mirror of the reference model ``src.models.sindy.SINDY`` (libs_m/ct/src/models/sindy.py) for the
cancer simulator: same constructor, ``fit`` / ``get_predictions`` / ``get_autoregressive_predictions`` and the
two RMSE methods it inherits from TimeVaryingCausalModel (time_varying_model.py:236-313), same attributes
(``joint_coefs``, ``global_equation_string``, ``feature_library_names``, ``feature_names``, ``insite`` ...).

What runs where
    fit                      K4 theta_gram + K5 population STLSQ (csrc/theta_gram.cu, fit_rollout.cu)
                             <- process_dataset_into_de_format + 4x pysindy SINDy.fit (sindy.py:160-213)
    population predictions   K6 ode_rollout <- _get_non_fine_tuned_predictions (sindy.py:371-431)
    INSITE predictions       K7 insite_bfgs (reference estimator) or K5b stlsq_batched (north-star estimator)
                             + K6 with per-row coefficients <- _get_fine_tuned_predictions (sindy.py:433-715)
    tau-step slicing         get_autoregressive_predictions (sindy.py:717-760)
    metrics                  numpy on the host, formulas of time_varying_model.py:236-313

Options: joint_model (11-term library), sindy_quantize, use_smoothed_finite_difference (smoothing pre-pass ahead of
K4), ablation_more_complex_basis_functions (degree-4 library, population models: csrc/poly_library.cu).

Not implemented (outside the cancer_sim hot path): Weak-SINDy, EQ_4/EQ_5 datasets, smooth_input_data,
ray-tune finetune, the degree-4 library for the joint model or under INSITE.  They raise NotImplementedError.
"""
import logging

import numpy as np
import torch

from . import device as dev

logger = logging.getLogger(__name__)

FEATURE_LIBRARY_NAMES = ['1', 'x0', 'u0', 'x0 u0']      # PolynomialLibrary(degree=2, interaction_only=True)
FEATURE_NAMES = ['x0', 'u0']
# joint model (model.joint_model=True, treatment_mode='multilabel'): one ODE over [x0, chemo, radio, static]
JOINT_FEATURE_LIBRARY_NAMES = ['1', 'x0', 'u0', 'u1', 'u2', 'x0 u0', 'x0 u1', 'x0 u2', 'u0 u1', 'u0 u2', 'u1 u2']
JOINT_FEATURE_NAMES = ['x0', 'u0', 'u1', 'u2']


def equation_string_core(feature_library_names, feature_names, coefs, quantize=False, quantize_round_to=3):
    """convert_sindy_model_to_sympyjax_model_core (pkpd/utils.py:378-391): the logged equation format."""
    names = [fn.replace(' ', '*') for fn in feature_library_names]
    out = []
    for fln in names:
        for i in range(len(feature_names)):
            fln = fln.replace(f'x{i}', feature_names[i])
        out.append(fln)
    s = ''
    for i, coef in enumerate(coefs):
        if np.abs(coef) > 1e-3:
            if quantize:
                coef = np.round(coef, quantize_round_to)
            s += f'+{coef}*' + out[i]
    return s


class SINDY:
    model_type = 'sindy_regressor'
    tuning_criterion = 'rmse'

    def __init__(self, args, dataset_collection=None, autoregressive=None, has_vitals=None, **kwargs):
        self.dataset_collection = dataset_collection
        if dataset_collection is not None:
            self.autoregressive = dataset_collection.autoregressive
            self.has_vitals = dataset_collection.has_vitals
        else:
            self.autoregressive, self.has_vitals = autoregressive, has_vitals
        self.hparams = args
        m = args.model
        self.dim_treatments = m.dim_treatments
        self.dim_vitals = m.dim_vitals
        self.dim_static_features = m.dim_static_features
        self.dim_outcome = m.dim_outcomes
        self.lag_features = m.lag_features
        self.input_size = self.dim_treatments + self.dim_static_features
        self.input_size += self.dim_vitals if self.has_vitals else 0
        self.input_size += self.dim_outcome if self.autoregressive else 0
        self.output_size = self.dim_outcome
        self.dt = dev.STANDARD_DT
        self.insite_val_error_threshold = m.insite_val_error_threshold
        self.global_equation_string = ''
        self.sindy_threshold = m.sindy_threshold
        self.sindy_alpha = m.sindy_alpha
        self.smooth_input_data = m.smooth_input_data
        self.sindy_quantize = m.sindy_quantize
        self.sindy_quantize_global_model_round_to = m.sindy_quantize_global_model_round_to
        self.lam = m.lam
        self.joint_model = m.joint_model
        # INSITE's optimiser = jax.scipy.optimize.minimize(method='BFGS', tol=1e-12) (sindy.py:627), restated in K7 with
        # jax's own semantics: `tol` is not forwarded (gtol stays 1e-5, maxiter 200 * n), and a zoom that fails (signed
        # bracket width <= 1e-10, or 30 trials) ends BFGS with status 3.  The reference's CURRENT code replaces those
        # rows' coefficients by the population's (:628-631) -- the default here, which reproduces the joint-model log
        # of 2023-05-16 (results/ablation/one_ode/...txt:6) to 2e-6.  The main-table log of 2023-05-14
        # (results/2_main_table/final_with_insite.txt:2362) was written by the revision before that fallback existed (the
        # commented-out line :632, `res.x` always): insite_zoom_failure_fallback=False reproduces it to 5e-15.
        self.insite_line_search = str(m.get('insite_line_search', 'jax'))
        # the robust line search reports status 3 for "exhausted at the noise floor, best point kept": no fallback there
        self.zoom_failure_fallback = bool(m.get('insite_zoom_failure_fallback', self.insite_line_search == 'jax'))
        self.insite = m.insite
        self.wsindy = m.wsindy
        self.use_smoothed_finite_difference = m.use_smoothed_finite_difference
        self.dataset_name = str(m.dataset_name).upper()
        self.ablation_more_complex_basis_functions = m.ablation_more_complex_basis_functions
        self.insight_recover_parametric_dist = m.insight_recover_parametric_dist
        self.treatment_mode = args.dataset.treatment_mode
        self.dim_one_hot_treatments = self.dim_treatments
        # counterfactual test sets of this package carry their device-resident compact cohort: the two RMSE methods then
        # evaluate it directly (compact_eval.py; one fit per (patient, t), no dense rows) unless this is switched off
        self.compact_evaluation = bool(m.get('compact_evaluation', True))
        self.insite_gtol = m.get('insite_gtol', None)            # None = the line-search flavour's default
        self.insite_max_iter = m.get('insite_max_iter', None)
        self.individualisation = getattr(m, 'individualisation', 'bfgs_rollout')
        self.ridge_prior_lam = getattr(m, 'ridge_prior_lam', 1e4)
        self.last_fit_info = {}
        if self.dataset_name != 'CANCER_SIM':
            raise NotImplementedError(f"dataset {m.dataset_name!r}: only cancer_sim is on the accelerated path")
        for flag in ('wsindy', 'smooth_input_data'):
            if getattr(self, flag):
                raise NotImplementedError(f"model.{flag}=True is outside the accelerated INSITE path (SURVEY.md §8a/f)")
        if self.ablation_more_complex_basis_functions and (self.joint_model or self.insite):
            raise NotImplementedError("ablation_more_complex_basis_functions: the degree-4 library is built for the four "
                                      "per-treatment population models (15 terms each); the joint model's 70 terms and "
                                      "the INSITE step over up to 60 coefficients are not")
        if self.joint_model:
            # the "one ODE" ablation (results/ablation/one_ode): 11-term library, multilabel treatments
            if self.treatment_mode != 'multilabel':
                raise NotImplementedError("model.joint_model=True is configured with treatment_mode='multilabel'")
            if self.insite:
                if m.get('individualisation', 'bfgs_rollout') != 'bfgs_rollout':
                    raise NotImplementedError("the joint model is individualised by the reference's BFGS estimator only")
        elif self.treatment_mode != 'multiclass':
            raise NotImplementedError("treatment_mode must be 'multiclass' for the per-treatment SINDy models")

    @staticmethod
    def set_hparams(model_args, new_args, input_size, model_type):
        model_args.lam = new_args['lam']

    def prepare_data(self):
        if self.dataset_collection is not None and not self.dataset_collection.processed_data_multi:
            self.dataset_collection.process_data_multi()

    def finetune(self, resources_per_trial=None, args=None):
        raise NotImplementedError("ray-tune hyper-parameter search is disabled in the reference (run.py:193)")

    # -- helpers -------------------------------------------------------------------------------------
    def _unscaled_inputs(self, dataset):
        """Un-scaling exactly as the reference does it (sindy.py:387-399, 554-567)."""
        sp = dataset.scaling_params
        prev = np.squeeze(dataset.data['prev_outputs'] * sp['output_stds'] + sp['output_means'], axis=-1)
        lo, hi = self.dim_outcome, self.dim_outcome + self.dim_static_features
        static = dataset.data['static_features'] * sp['inputs_stds'][lo:hi] + sp['input_means'][lo:hi]
        if self.treatment_mode == 'multilabel':      # (chemo, radio) applications -> code chemo + 2*radio
            cti = dataset.data['current_treatments'].astype(np.int64)                # sindy.py:396
            codes = (cti[..., 0] + 2 * cti[..., 1]).astype(np.uint8)
        else:
            # argmax of the one-hot rows (sindy.py:397).  The product's own dataset classes keep the index array the
            # one-hot encoding is built from (and drop it when 'current_treatments' is reassigned), so neither the
            # 1.2 GB one-hot array of a 10k/1k/1k collection nor its argmax (0.4 s) is needed here
            cached = getattr(dataset, 'treatment_codes_', None)
            codes_fn = getattr(dataset, '_treatment_codes', None)
            if cached is None and codes_fn is not None:
                cached = codes_fn()
            if cached is not None and cached.shape == prev.shape:
                codes = cached
            else:
                codes = np.argmax(dataset.data['current_treatments'], axis=-1).astype(np.uint8)
        seq = dataset.data['sequence_lengths'].astype(np.int64)
        return prev, static[:, 0], codes, seq

    # -- fit (sindy.py:145-338) ----------------------------------------------------------------------
    def fit(self, train_f, val_f=None):
        self.prepare_data()
        dev.require_cuda()
        sp = train_f.scaling_params
        prev, static, codes, seq = self._unscaled_inputs(train_f)
        unscaled_outputs = np.squeeze(train_f.data['unscaled_outputs'], axis=-1)
        # pkpd/utils.py:554: 60-long volume = [prev_outputs[:,0] | unscaled_outputs]
        vol = np.concatenate((prev[:, 0].reshape(-1, 1), unscaled_outputs), axis=1)
        n, T = vol.shape
        chemo = np.zeros((n, T)); radio = np.zeros((n, T))
        chemo[:, :T - 1] = (codes & 1)
        radio[:, :T - 1] = (codes >> 1) & 1
        vol_d, chemo_d, radio_d = dev.to_device(vol), dev.to_device(chemo), dev.to_device(radio)
        seq_d = dev.to_device(seq.astype(np.float64))
        if self.use_smoothed_finite_difference:
            # SmoothedFiniteDifference(window_length=2, polyorder=1), sindy.py:196-198: the trajectories are smoothed,
            # then differentiated and expanded as usual
            vol_d = dev.smooth_snippets(vol_d, chemo_d, radio_d, seq_d, joint=bool(self.joint_model))
        if self.ablation_more_complex_basis_functions:
            # PolynomialLibrary(degree=4, interaction_only=False), sindy.py:185-186: R factors instead of normal equations
            rf = dev.poly_tsqr(vol_d, chemo_d, radio_d, seq_d, dev.to_device(static), fd_dt=self.dt)
            coefs, support = dev.poly_stlsq(rf, threshold=self.sindy_threshold, alpha=self.sindy_alpha, max_iter=100)
            torch.cuda.current_stream().synchronize()
            self.joint_coefs = coefs.cpu().numpy()                             # (4, 15), sindy.py:334
            self.support_ = support.cpu().numpy().astype(bool)
            self.population_stats_ = rf.cpu().numpy().copy()
            self.feature_library_names = list(dev.POLY_FEATURE_LIBRARY_NAMES)
            self.feature_names = list(FEATURE_NAMES)
            strs = [equation_string_core(self.feature_library_names, self.feature_names, self.joint_coefs[a],
                                         quantize=self.sindy_quantize,
                                         quantize_round_to=self.sindy_quantize_global_model_round_to) for a in range(4)]
            self.global_equation_string = (f'Treatment 0: x_dot = {strs[0]} | Treatment 1: x_dot = {strs[1]} | '
                                           f'Treatment 2: x_dot = {strs[2]} | Treatment 3: x_dot = {strs[3]}')
            logger.info('[Model]: ' + self.global_equation_string)
            return self
        stats = dev.theta_gram(vol_d, chemo_d, radio_d, seq_d, dev.to_device(static), fd_dt=self.dt,
                               joint=bool(self.joint_model))
        if self.joint_model:
            coefs11, support11, c44 = dev.stlsq_joint(stats, threshold=self.sindy_threshold, alpha=self.sindy_alpha,
                                                      max_iter=100, drop_below=1e-3)
            torch.cuda.current_stream().synchronize()
            self.joint_coefs = coefs11.cpu().numpy()[None, :]                  # (1, 11), sindy.py:336
            self.support_ = support11.cpu().numpy().astype(bool)[None, :]
            self.rollout_coefs_ = c44.cpu().numpy()                            # thresholded expression per treatment
            self.population_stats_ = stats.cpu().numpy().copy()
            self.feature_library_names = list(JOINT_FEATURE_LIBRARY_NAMES)
            self.feature_names = list(JOINT_FEATURE_NAMES)
            str_0 = equation_string_core(self.feature_library_names, self.feature_names, self.joint_coefs[0],
                                         quantize=self.sindy_quantize,
                                         quantize_round_to=self.sindy_quantize_global_model_round_to)
            self.global_equation_string = f'Joint Model: x_dot = {str_0}'      # sindy.py:314
            logger.info('[Model Raw]: ' + self.global_equation_string)
            return self
        coefs, support = dev.stlsq_population(stats, threshold=self.sindy_threshold, alpha=self.sindy_alpha, max_iter=100)
        torch.cuda.current_stream().synchronize()
        self.joint_coefs = coefs.cpu().numpy()
        self.support_ = support.cpu().numpy().astype(bool)
        self.population_stats_ = stats.cpu().numpy().copy()
        self.feature_library_names = list(FEATURE_LIBRARY_NAMES)
        self.feature_names = list(FEATURE_NAMES)
        strs = [equation_string_core(self.feature_library_names, self.feature_names, self.joint_coefs[a],
                                     quantize=self.sindy_quantize,
                                     quantize_round_to=self.sindy_quantize_global_model_round_to) for a in range(4)]
        self.global_equation_string = (f'Treatment 0: x_dot = {strs[0]} | Treatment 1: x_dot = {strs[1]} | '
                                       f'Treatment 2: x_dot = {strs[2]} | Treatment 3: x_dot = {strs[3]}')
        logger.info('[Model]: ' + self.global_equation_string)
        return self

    def _population_rollout_coefs(self):
        """(4,4) coefficients of the expression the reference integrates for population predictions: terms with
        |c| > 1e-3, rounded to sindy_quantize_global_model_round_to decimals when sindy_quantize is set
        (convert_sindy_model_to_sympyjax_model_core, pkpd/utils.py:386-391); the joint model's 11 terms restricted to each
        treatment code.  The kernels take them with drop_below < 0 (nothing else to drop)."""
        c = np.asarray(self.joint_coefs, dtype=np.float64)
        e = np.where(np.abs(c) > 1e-3, np.round(c, self.sindy_quantize_global_model_round_to) if self.sindy_quantize else c, 0.0)
        if self.joint_model:
            return (_JOINT_TO_PER_TREATMENT @ e[0]).reshape(4, 4)
        return e

    # -- predictions ---------------------------------------------------------------------------------
    def get_predictions(self, dataset):
        if not self.insite:
            predictions = self._get_non_fine_tuned_predictions(dataset)
        else:
            predictions = self._get_fine_tuned_predictions(dataset)
        assert not np.any(np.isnan(predictions)), 'Predictions contains NaN'
        return predictions

    def _rollout(self, prev, static, codes, coefs_dev, drop_below):
        if self.ablation_more_complex_basis_functions:
            pred = dev.poly_rollout(dev.to_device(np.ascontiguousarray(prev[:, 0])), dev.to_device(static),
                                    dev.to_device(codes, dtype=torch.uint8), coefs_dev, dt=self.dt,
                                    substeps=dev.STEPS_FOR_DT, drop_below=drop_below)
            torch.cuda.current_stream().synchronize()
            return pred.cpu().numpy()
        pred = dev.ode_rollout(dev.to_device(np.ascontiguousarray(prev[:, 0])), dev.to_device(static),
                               dev.to_device(codes, dtype=torch.uint8), coefs_dev, dt=self.dt,
                               substeps=dev.STEPS_FOR_DT, drop_below=drop_below)
        torch.cuda.current_stream().synchronize()
        return pred.cpu().numpy()

    def _get_non_fine_tuned_predictions(self, dataset):
        """Open-loop rollout of the population ODE (terms with |c| <= 1e-3 dropped, pkpd/utils.py:388)."""
        sp = dataset.scaling_params
        prev, static, codes, _ = self._unscaled_inputs(dataset)
        un = self._rollout(prev, static, codes, dev.to_device(self._population_rollout_coefs()), -1.0)
        return ((un - sp['output_means']) / sp['output_stds'])[..., None]

    def individualised_coefficients(self, dataset, projection_horizon=1):
        """Per-row coefficient matrices (R,4,4) on the device + diagnostics."""
        prev, static, codes, seq = self._unscaled_inputs(dataset)
        x = dev.to_device(prev)
        cd = dev.to_device(codes, dtype=torch.uint8)
        st = dev.to_device(static)
        theta0 = dev.to_device(self.joint_coefs)
        W = prev.shape[1]
        if self.individualisation == 'bfgs_rollout':
            coefs, status, fval = dev.insite_bfgs(x, cd, dev.to_device(seq, dtype=torch.int32), projection_horizon,
                                                  st, theta0, lam=self.lam, gtol=self.insite_gtol, max_iter=self.insite_max_iter,
                                                  joint=bool(self.joint_model), line_search=self.insite_line_search)
            if self.zoom_failure_fallback:
                # sindy.py:628-631: "if zoom fails, fall back to default value" (res.status == 3 -> population coefficients)
                failed = (status & 255) == 3
                coefs[failed] = theta0.reshape(coefs.shape[1:])
            if self.joint_model:
                # per-row 11-term ODE restricted to each treatment code: the 4-term form the rollout kernel integrates
                coefs = torch.matmul(coefs, dev.to_device(_JOINT_TO_PER_TREATMENT).T).reshape(-1, 4, 4).contiguous()
            torch.cuda.current_stream().synchronize()
            st_np = status.cpu().numpy()
            self.last_fit_info = {'estimator': 'bfgs_rollout', 'status_low_byte': np.bincount(st_np[st_np >= 0] & 0xff, minlength=8),
                                  'skipped': int((st_np < 0).sum()), 'iterations_mean': float((st_np[st_np >= 0] >> 8).mean()) if (st_np >= 0).any() else 0.0,
                                  'objective_start_mean': float(fval[:, 0].mean().item()),
                                  'objective_end_mean': float(fval[:, 1].mean().item())}
        elif self.individualisation == 'ridge_prior_stlsq':
            fit_len = np.clip(seq - projection_horizon, 0, W - 1).astype(np.int32)
            coefs = dev.stlsq_batched(x, cd, dev.to_device(fit_len, dtype=torch.int32), st, theta0,
                                      lam=self.ridge_prior_lam, threshold=self.sindy_threshold, support_tol=1e-3,
                                      fd_dt=self.dt)
            self.last_fit_info = {'estimator': 'ridge_prior_stlsq', 'lam': self.ridge_prior_lam}
        else:
            raise ValueError(f"unknown individualisation estimator {self.individualisation!r}")
        return coefs, (prev, static, codes, seq)

    def _get_fine_tuned_predictions(self, dataset, projection_horizon=1):
        """INSITE: individualise per row, then roll out with the row's own coefficients (no 1e-3 term filter,
        as predict_with_reduced_coefs, sindy.py:658)."""
        sp = dataset.scaling_params
        coefs, (prev, static, codes, _) = self.individualised_coefficients(dataset, projection_horizon)
        un = self._rollout(prev, static, codes, coefs, -1.0)
        scaled = (un - sp['output_means']) / sp['output_stds']
        assert not np.any(np.isnan(scaled) | np.isinf(scaled)), 'Scaled_preds contains NaN or Inf'
        return scaled[..., None]

    def get_autoregressive_predictions(self, dataset):
        H = self.hparams.dataset.projection_horizon
        if not self.insite:
            scaled = self._get_non_fine_tuned_predictions(dataset)
        else:
            scaled = self._get_fine_tuned_predictions(dataset, projection_horizon=H)
        assert scaled.ndim == 3 and scaled.shape[2] == 1
        seq = dataset.data['sequence_lengths'].astype(np.int64)
        R, W, _ = scaled.shape
        # dynamic_slice(preds, (i, max(1, sl-H), 0), (1,H,1)) -- start clamped so the slice fits (sindy.py:729-733)
        lo = np.clip(np.maximum(1, seq - H), 0, W - H)
        idx = lo[:, None] + np.arange(H)[None, :]
        return scaled[np.arange(R)[:, None], idx, :]

    # -- metrics (time_varying_model.py:236-313) -------------------------------------------------------
    def _compact_metrics(self, dataset, kind):
        """The RMSEs of a counterfactual test set from its compact cohort (None when the dense path has to run: no
        compact cohort, switched off, or INSITE on the joint model)."""
        compact = getattr(dataset, 'compact_', None) if self.compact_evaluation else None
        if compact is None or compact[0].kind != kind or (self.joint_model and self.insite) or \
                self.ablation_more_complex_basis_functions:     # the compact kernels integrate the affine ODE
            return None
        from . import compact_eval as ce
        cohort, static = compact
        unscale = self.hparams.exp.unscale_rmse
        scale = 1.0 if unscale else float(dataset.scaling_params['output_stds'])
        theta0 = dev.to_device(self.joint_coefs)       # not reached for the joint model with INSITE
        if self.insite:
            coefs, diag = ce.individualise(cohort, static, theta0, self.individualisation, self.lam, self.ridge_prior_lam,
                                           self.sindy_threshold, self.zoom_failure_fallback, self.dt, gtol=self.insite_gtol,
                                           max_iter=self.insite_max_iter, line_search=self.insite_line_search)
            if 'status' in diag:
                st_np = diag['status'].cpu().numpy()
                ok = st_np >= 0
                self.last_fit_info = {'estimator': 'bfgs_rollout', 'fits': int(ok.sum()), 'rows_covered': int(cohort.total_rows),
                                      'status_low_byte': np.bincount(st_np[ok] & 0xff, minlength=8),
                                      'iterations_mean': float((st_np[ok] >> 8).mean()) if ok.any() else 0.0}
            drop = -1.0
        else:
            coefs, drop = dev.to_device(self._population_rollout_coefs()), -1.0
        sums = ce._finish(ce.evaluate(cohort, static, coefs, drop, self.dt))
        pct = self.hparams.exp.percentage_rmse
        if kind == 'one_step':
            return ce.one_step_rmses(sums, cohort.T - 1, dataset.norm_const, pct, scale)
        return ce.n_step_rmses(sums, cohort.H, dataset.norm_const, pct, scale)

    def get_normalised_masked_rmse(self, dataset, one_step_counterfactual=False):
        if one_step_counterfactual:
            fast = self._compact_metrics(dataset, 'one_step')
            if fast is not None:
                return fast
        outputs_scaled = self.get_predictions(dataset)
        unscale = self.hparams.exp.unscale_rmse
        percentage = self.hparams.exp.percentage_rmse
        act = dataset.data['active_entries']
        if unscale:
            sp = dataset.scaling_params
            err = outputs_scaled * sp['output_stds'] + sp['output_means'] - dataset.data['unscaled_outputs']
        else:
            err = outputs_scaled - dataset.data['outputs']
        mse = (err ** 2) * act
        mse_orig = (mse.sum(0).sum(-1) / act.sum(0).sum(-1)).mean()
        rmse_orig = np.sqrt(mse_orig) / dataset.norm_const
        rmse_all = np.sqrt(mse.sum() / act.sum()) / dataset.norm_const
        if percentage:
            rmse_orig *= 100.0
            rmse_all *= 100.0
        if one_step_counterfactual:
            n, _, od = act.shape
            last = act - np.concatenate([act[:, 1:, :], np.zeros((n, 1, od))], axis=1)
            mse_last = ((err ** 2) * last).sum() / last.sum()
            rmse_last = np.sqrt(mse_last) / dataset.norm_const
            if percentage:
                rmse_last *= 100.0
            return rmse_orig, rmse_all, rmse_last
        return rmse_orig, rmse_all

    def get_normalised_n_step_rmses(self, dataset, datasets_mc=None):
        assert hasattr(dataset, 'data_processed_seq')
        if datasets_mc is None:
            fast = self._compact_metrics(dataset, 'treatment_seq')
            if fast is not None:
                return fast
        unscale = self.hparams.exp.unscale_rmse
        percentage = self.hparams.exp.percentage_rmse
        outputs_scaled = self.get_autoregressive_predictions(dataset if datasets_mc is None else datasets_mc)
        seq = dataset.data_processed_seq
        if unscale:
            sp = dataset.scaling_params
            mse = ((outputs_scaled * sp['output_stds'] + sp['output_means'] - seq['unscaled_outputs']) ** 2) \
                * seq['active_entries']
        else:
            mse = ((outputs_scaled - seq['outputs']) ** 2) * seq['active_entries']
        nan_idx = np.unique(np.where(np.isnan(seq['outputs']))[0])
        not_nan = np.setdiff1d(np.arange(outputs_scaled.shape[0]), nan_idx)
        mse_orig = mse[not_nan].sum(0).sum(-1) / seq['active_entries'][not_nan].sum(0).sum(-1)
        rmses = np.sqrt(mse_orig) / dataset.norm_const
        if percentage:
            rmses *= 100.0
        return rmses


def _joint_to_per_treatment_matrix():
    """(16, 11): row 4*code + m = coefficient of [1, x0, u2, x0*u2][m] under treatment code chemo + 2*radio as a
    linear function of the 11 joint coefficients [1 x0 u0 u1 u2 x0u0 x0u1 x0u2 u0u1 u0u2 u1u2] (u0 chemo, u1 radio)."""
    M = np.zeros((16, 11))
    for code in range(4):
        ch, ra = float(code & 1), float(code >> 1)
        M[4 * code + 0, [0, 2, 3, 8]] = [1.0, ch, ra, ch * ra]
        M[4 * code + 1, [1, 5, 6]] = [1.0, ch, ra]
        M[4 * code + 2, [4, 9, 10]] = [1.0, ch, ra]
        M[4 * code + 3, 7] = 1.0
    return M


_JOINT_TO_PER_TREATMENT = _joint_to_per_treatment_matrix()


def run_experiment(args, dataset_collection=None):
    """Body of runnables/train_sindy.py::main (:38-113) without hydra / MLflow: returns the same result dict."""
    from .dataset import SyntheticCancerDatasetCollection
    if dataset_collection is None:
        d = args.dataset
        dataset_collection = SyntheticCancerDatasetCollection(
            d.chemo_coeff, d.radio_coeff, {'train': d.num_patients.train, 'val': d.num_patients.val,
                                           'test': d.num_patients.test}, seed=d.seed, window_size=d.window_size,
            max_seq_length=d.max_seq_length, projection_horizon=d.projection_horizon, lag=d.lag,
            cf_seq_mode=d.cf_seq_mode, treatment_mode=d.treatment_mode)
    dataset_collection.process_data_multi()
    model = SINDY(args, dataset_collection)
    model.fit(dataset_collection.train_f, dataset_collection.val_f)
    results = {}
    orig, all_, last = model.get_normalised_masked_rmse(dataset_collection.test_cf_one_step, one_step_counterfactual=True)
    results.update({'encoder_test_rmse_all': all_, 'encoder_test_rmse_orig': orig, 'encoder_test_rmse_last': last})
    rmses = model.get_normalised_n_step_rmses(dataset_collection.test_cf_treatment_seq)
    results.update({f'decoder_test_rmse_{k + 2}-step': v for k, v in enumerate(rmses)})
    results.update({'global_equation_string': model.global_equation_string, 'fine_tuned': model.insite})
    return results, model
