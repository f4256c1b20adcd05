"""Host mirror of the reference dataset layer for the cancer simulator
(libs_m/ct/src/data/cancer_sim/dataset.py, libs_m/ct/src/data/dataset_collection.py):

    SyntheticCancerDataset            dataset.py:17-552   (generation via the CUDA simulators,
                                                           process_data :92-192,
                                                           process_sequential_test :395-473,
                                                           process_sequential_multi :533-552)
    SyntheticCancerDatasetCollection  dataset.py:555-605  (+ process_data_multi,
                                                           dataset_collection.py:74-86)

Same constructor arguments, attributes and ``.data`` dictionaries (keys, shapes, dtypes: SURVEY.md
App. D), so ``SINDY`` and the reference's ``train_sindy.main`` can consume them unchanged.  The
reference's O(R*T) python loops (one-hot encoding :131-141, active mask :162-164, per-row slicing
:427-439) are replaced by vectorised numpy; the values are identical.  After the simulators these transforms are
what a default-size experiment spends its time in (0.5 of 0.66 s), so they avoid strided writes and temporaries:
one-hot rows are gathered from an identity table, shifted copies are written into preallocated arrays, and the
reference's two deepcopies of the whole dictionary (:470, :541) become dictionary copies that share the arrays
(nothing on this path writes into them).
"""
import logging

import numpy as np

_EYE4 = np.vstack([np.eye(4), np.zeros((1, 4))])   # row 4 = all zero: a treatment pair that is not 0/1-valued


from .lazydict import LazyDict
from .cancer_simulation import (TUMOUR_DEATH_THRESHOLD, generate_params, get_scaling_params, simulate_factual,
                                simulate_counterfactual_1_step, simulate_counterfactuals_treatment_seq)

logger = logging.getLogger(__name__)


class SyntheticCancerDataset:
    """Tumour-growth simulator dataset (torch Dataset protocol: __len__ / __getitem__)."""

    def __init__(self, chemo_coeff, radio_coeff, num_patients, window_size, seq_length, subset_name,
                 mode='factual', projection_horizon=None, seed=None, lag=0, cf_seq_mode='sliding_treatment',
                 treatment_mode='multiclass'):
        if seed is not None:
            np.random.seed(seed)
        self.chemo_coeff = chemo_coeff
        self.radio_coeff = radio_coeff
        self.window_size = window_size
        self.num_patients = num_patients
        self.params = generate_params(num_patients, chemo_coeff=chemo_coeff, radio_coeff=radio_coeff,
                                      window_size=window_size, lag=lag)
        self.subset_name = subset_name
        if mode == 'factual':
            self.data = simulate_factual(self.params, seq_length)
        elif mode == 'counterfactual_one_step':
            # lazy: the dense (R, T) rows stay on the device in compact form until something reads them
            self.data = simulate_counterfactual_1_step(self.params, seq_length, lazy=True)
        elif mode == 'counterfactual_treatment_seq':
            assert projection_horizon is not None
            self.data = simulate_counterfactuals_treatment_seq(self.params, seq_length, projection_horizon, cf_seq_mode,
                                                               lazy=True)
        else:
            raise ValueError(f"unknown mode {mode!r}")
        self.processed = False
        self.processed_sequential = False
        self.processed_autoregressive = False
        self.treatment_mode = treatment_mode
        self.exploded = False
        self.norm_const = TUMOUR_DEATH_THRESHOLD

    def __getitem__(self, index):
        return {k: v[index] for k, v in self.data.items() if hasattr(v, '__len__') and len(v) == len(self)}

    def __len__(self):
        return self.data['sequence_lengths'].shape[0]   # = current_covariates.shape[0] (:86), without building that array

    @property
    def compact_(self):
        """(CompactCohort, static feature on the device) of a counterfactual test set whose dictionaries have not been
        tampered with, else None: what SINDY's RMSE methods evaluate instead of the dense rows."""
        d = self.data_original if (self.processed_sequential and not self.processed_autoregressive) else self.data
        attrs = getattr(d, 'attrs', None)
        return attrs.get('compact') if attrs else None

    def get_scaling_params(self):
        return get_scaling_params(self.data)

    # The reference caches whole collections with shelve (run_utils.py:4-19), i.e. pickles them: the row builders are
    # closures and stay behind; a restored dataset reads the (then materialised) arrays instead.
    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop('_treatment_rows', None)
        state.pop('_covariate_rows', None)
        state.pop('_treatment_codes', None)
        return state

    def _rows_of_treatments(self, rows, cols):
        f = getattr(self, '_treatment_rows', None)
        return f(rows, cols) if f is not None else self.data['current_treatments'][rows, cols, :]

    def _rows_of_covariates(self, rows, cols):
        f = getattr(self, '_covariate_rows', None)
        return f(rows, cols) if f is not None else self.data['current_covariates'][rows, cols, :]

    # -- dataset.py:92-192 ---------------------------------------------------------------------------
    def process_data(self, scaling_params, include_continuous_treatment=False):
        if self.processed:
            logger.info(f'{self.subset_name} Dataset already processed')
            return self.data
        mean, std = scaling_params
        mean['chemo_application'] = 0
        mean['radio_application'] = 0
        std['chemo_application'] = 1
        std['radio_application'] = 1
        cols = ['cancer_volume', 'patient_types', 'chemo_application', 'radio_application']
        input_means = mean[cols].values.flatten()
        input_stds = std[cols].values.flatten()
        base = self.data
        # counterfactual test sets arrive with their dense rows pending (counterfactual._lazy_dense): then every array
        # derived from them is pending too, and reassigning any simulator output drops the compact cohort
        lazy_big = isinstance(base, LazyDict) and base.pending('cancer_volume')
        data = base.copy() if isinstance(base, LazyDict) else LazyDict(base)
        mv, sv = mean['cancer_volume'], std['cancer_volume']
        patient_types = np.asarray((base['patient_types'] - mean['patient_types']) / std['patient_types'])
        seq_len = base['sequence_lengths']
        R = seq_len.shape[0]
        cache = {}

        def scaled_volume():
            if 'cv' not in cache:
                cache['cv'] = (data['cancer_volume'] - mv) / sv
            return cache['cv']

        def width():
            return data['cancer_volume'].shape[1]

        def put(key, fn):
            if lazy_big:
                data.set_lazy(key, fn)
            else:
                data[key] = fn()

        if self.treatment_mode == 'multiclass':
            # one_hot[..., a] = (c, r) == ((0,0), (1,0), (0,1), (1,1))[a]  (:131-141): a row of the identity table
            def treatment_index():
                if 'idx' not in cache:
                    c, r = data['chemo_application'][:, :-1], data['radio_application'][:, :-1]
                    idx = (c == 1).astype(np.int8) + 2 * (r == 1).astype(np.int8)
                    bad = ~(((c == 0) | (c == 1)) & ((r == 0) | (r == 1)))
                    if bad.any():
                        idx[bad] = 4
                    cache['idx'] = idx
                return cache['idx']
            self._treatment_rows = lambda rows, cols: _EYE4[treatment_index()[rows, cols]]   # current_treatments[rows, cols, :]
            data.set_lazy('current_treatments', lambda: _EYE4[treatment_index()])
            # np.argmax(one_hot, -1) without the reduction (an all-zero row has argmax 0): read by SINDY
            self._treatment_codes = lambda: np.where(treatment_index() == 4, 0, treatment_index()).astype(np.uint8)
            self.treatment_codes_ = None if lazy_big else self._treatment_codes()
            data.on_set('current_treatments', lambda: (setattr(self, 'treatment_codes_', None),
                                                       setattr(self, '_treatment_codes', None)))
            k_tr = 4
        elif self.treatment_mode == 'multilabel':
            def treatments():
                chemo, radio = data['chemo_application'], data['radio_application']
                t = np.empty((R, width() - 1, 2))
                t[..., 0] = chemo[:, :-1]
                t[..., 1] = radio[:, :-1]
                return t
            self._treatment_rows = lambda rows, cols: np.stack([data['chemo_application'][rows, cols],
                                                                data['radio_application'][rows, cols]], axis=-1)
            data.set_lazy('current_treatments', treatments)
            k_tr = 2
        else:
            raise ValueError(self.treatment_mode)

        def prev_treatments():   # zero row for t = 0, then current_treatments[:, :-1] (:141, :183-185)
            p = np.zeros((R, width() - 1, k_tr))
            p[:, 1:, :] = data['current_treatments'][:, :-1, :]
            return p
        data.set_lazy('prev_treatments', prev_treatments)

        def current_covariates():
            cov = np.empty((R, width() - 1, 2))
            cov[..., 0] = scaled_volume()[:, :-1]
            cov[..., 1] = patient_types[:, None]
            return cov
        data.set_lazy('current_covariates', current_covariates)
        self._covariate_rows = lambda rows, cols: np.stack(
            [scaled_volume()[rows, cols], np.broadcast_to(patient_types[rows], np.shape(cols))], axis=-1)

        output_means = mean[['cancer_volume']].values.flatten()[0]
        output_stds = std[['cancer_volume']].values.flatten()[0]
        put('outputs', lambda: scaled_volume()[:, 1:, np.newaxis])
        put('active_entries', lambda: (np.arange(width() - 1)[None, :] < seq_len.astype(np.int64)[:, None])
            .astype(np.float64)[:, :, None])
        put('unscaled_outputs', lambda: data['outputs'] * sv + mv)
        self.scaling_params = {'input_means': input_means, 'inputs_stds': input_stds,
                               'output_means': output_means, 'output_stds': output_stds}
        # = current_covariates[:, :, :1] and current_covariates[:, 0, 1:] of the reference (:186-187), same values
        put('prev_outputs', lambda: scaled_volume()[:, :-1, np.newaxis])
        data['static_features'] = patient_types[:, np.newaxis].copy()
        if lazy_big:     # from here on, assigning any of these keys makes the compact cohort stale: it is dropped
            data.watch_attrs(('cancer_volume', 'chemo_application', 'radio_application', 'sequence_lengths', 'patient_types',
                              'prev_outputs', 'outputs', 'unscaled_outputs', 'current_treatments', 'static_features',
                              'active_entries'))
        self.data = data
        self.processed = True
        return self.data

    # -- dataset.py:395-473 (encoder_r=None path) ------------------------------------------------------
    def process_sequential_test(self, projection_horizon, encoder_r=None, save_encoder_r=False,
                                include_continuous_treatment=False):
        assert self.processed
        if encoder_r is not None:
            raise NotImplementedError("encoder representations belong to the neural baselines (out of scope)")
        if self.processed_sequential:
            return self.data
        H = projection_horizon
        full = self.data
        seq_len = full['sequence_lengths'].astype(np.int64)
        R = seq_len.shape[0]
        fact = seq_len - H
        scaling = self.scaling_params

        def build():
            outputs = full['outputs']
            W = outputs.shape[1]
            rows = np.arange(R)[:, None]
            k = np.arange(H)[None, :]
            # prev = prev_treatments[:, 1:] = current_treatments[:, :-1]  (W - 1 entries)
            prev_idx = np.mod(fact[:, None] - 1 + k, W - 1)   # python slices with a negative start never occur (sl > H)
            cov_last = self._rows_of_covariates(np.arange(R), fact - 1)                    # current_covariates[i, fact-1]
            out = {
                'active_encoder_r': (np.arange(W - H)[None, :] < fact[:, None]).astype(np.float64),
                'prev_treatments': self._rows_of_treatments(rows, prev_idx),
                'current_treatments': self._rows_of_treatments(rows, fact[:, None] + k),
                'current_covariates': np.repeat(cov_last[:, None, :], H, axis=1),
                'outputs': outputs[rows, fact[:, None] + k, :],
            }
            out['prev_outputs'] = out['current_covariates'][:, :, :1]
            out['static_features'] = out['current_covariates'][:, 0, 1:]
            out['unscaled_outputs'] = out['outputs'] * scaling['output_stds'] + scaling['output_means']
            return out
        group = ('active_encoder_r', 'prev_treatments', 'current_treatments', 'current_covariates', 'outputs',
                 'prev_outputs', 'static_features', 'unscaled_outputs')
        seq = LazyDict()
        if isinstance(full, LazyDict) and full.pending('cancer_volume'):
            seq.set_lazy_group(group, build)      # built when somebody reads the sliced rows (the compact path does not)
        else:
            seq.update(build())
        seq['sequence_lengths'] = np.full(R, float(H))
        seq['active_entries'] = np.ones((R, H, 1))
        seq['patient_types'] = full['patient_types']
        seq['patient_ids_all_trajectories'] = full['patient_ids_all_trajectories']
        seq['patient_current_t'] = full['patient_current_t']
        self.data_original = self.data.copy()    # the reference deep-copies (:470); arrays are shared here
        self.data = seq
        self.processed_sequential = True
        return self.data

    # -- dataset.py:533-552 --------------------------------------------------------------------------
    def process_sequential_multi(self, projection_horizon, include_continuous_treatment=False):
        assert self.processed_sequential
        if not self.processed_autoregressive:
            self.data_processed_seq = self.data
            self.data = self.data_original.copy()  # (:541)
            self.data['future_past_split'] = self.data['sequence_lengths'] - projection_horizon
            self.processed_autoregressive = True
        return self.data

    def explode_trajectories(self, projection_horizon):
        raise NotImplementedError("explode_trajectories feeds the encoder-decoder baselines (dataset.py:194-280); "
                                  "it is not on the SINDy/INSITE path")


class SyntheticCancerDatasetCollection:
    """train_f, val_f, test_cf_one_step, test_cf_treatment_seq -- one RNG stream, seeded once
    (dataset.py:589), subsets generated in the reference's order (:591-601)."""

    def __init__(self, chemo_coeff, radio_coeff, num_patients, seed=100, window_size=15, max_seq_length=60,
                 projection_horizon=5, lag=0, cf_seq_mode='sliding_treatment', treatment_mode='multiclass', **kwargs):
        self.seed = seed
        self.processed_data_encoder = False
        self.processed_data_decoder = False
        self.processed_data_multi = False
        self.processed_data_msm = False
        np.random.seed(seed)
        self.train_f = SyntheticCancerDataset(chemo_coeff, radio_coeff, num_patients['train'], window_size,
                                              max_seq_length, 'train', lag=lag, treatment_mode=treatment_mode)
        self.val_f = SyntheticCancerDataset(chemo_coeff, radio_coeff, num_patients['val'], window_size,
                                            max_seq_length, 'val', lag=lag, treatment_mode=treatment_mode)
        self.test_cf_one_step = SyntheticCancerDataset(chemo_coeff, radio_coeff, num_patients['test'], window_size,
                                                       max_seq_length, 'test', mode='counterfactual_one_step',
                                                       lag=lag, treatment_mode=treatment_mode)
        self.test_cf_treatment_seq = SyntheticCancerDataset(chemo_coeff, radio_coeff, num_patients['test'],
                                                            window_size, max_seq_length, 'test',
                                                            mode='counterfactual_treatment_seq',
                                                            projection_horizon=projection_horizon, lag=lag,
                                                            cf_seq_mode=cf_seq_mode, treatment_mode=treatment_mode)
        self.projection_horizon = projection_horizon
        self.autoregressive = True
        self.has_vitals = False
        self.train_scaling_params = self.train_f.get_scaling_params()

    def process_data_multi(self, include_continuous_treatment=False):
        """dataset_collection.py:74-86."""
        self.train_f.process_data(self.train_scaling_params)
        if getattr(self, 'val_f', None) is not None:
            self.val_f.process_data(self.train_scaling_params)
        self.test_cf_one_step.process_data(self.train_scaling_params)
        self.test_cf_treatment_seq.process_data(self.train_scaling_params)
        self.test_cf_treatment_seq.process_sequential_test(self.projection_horizon)
        self.test_cf_treatment_seq.process_sequential_multi(self.projection_horizon)
        self.processed_data_multi = True
