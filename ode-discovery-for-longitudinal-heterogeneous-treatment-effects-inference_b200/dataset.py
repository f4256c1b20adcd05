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
            self.data = simulate_counterfactual_1_step(self.params, seq_length)
        elif mode == 'counterfactual_treatment_seq':
            assert projection_horizon is not None
            self.data = simulate_counterfactuals_treatment_seq(self.params, seq_length, projection_horizon, cf_seq_mode)
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
        return self.data['current_covariates'].shape[0]

    def get_scaling_params(self):
        return get_scaling_params(self.data)

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

        cancer_volume = (self.data['cancer_volume'] - mean['cancer_volume']) / std['cancer_volume']
        patient_types = (self.data['patient_types'] - mean['patient_types']) / std['patient_types']
        width = cancer_volume.shape[1]
        patient_types = np.asarray(patient_types)

        chemo = self.data['chemo_application']
        radio = self.data['radio_application']
        seq_len = self.data['sequence_lengths']
        R = chemo.shape[0]
        if self.treatment_mode == 'multiclass':
            # one_hot[..., a] = (c, r) == ((0,0), (1,0), (0,1), (1,1))[a]  (:131-141): a row of the identity table
            c, r = chemo[:, :-1], radio[:, :-1]
            idx = (c == 1).astype(np.int8) + 2 * (r == 1).astype(np.int8)
            bad = ~(((c == 0) | (c == 1)) & ((r == 0) | (r == 1)))
            if bad.any():
                idx[bad] = 4
            one_hot = _EYE4[idx]
            # np.argmax(one_hot, -1) without the reduction (an all-zero row has argmax 0): read by SINDY
            self.treatment_codes_ = np.where(idx == 4, 0, idx).astype(np.uint8)
            self._codes_owner = one_hot
            prev_tr = np.zeros((R, width - 1, 4))
            prev_tr[:, 1:, :] = one_hot[:, :-1, :]
            self.data['current_treatments'] = one_hot
        elif self.treatment_mode == 'multilabel':
            treatments = np.empty((R, width - 1, 2))
            treatments[..., 0] = chemo[:, :-1]
            treatments[..., 1] = radio[:, :-1]
            prev_tr = np.zeros((R, width - 1, 2))
            prev_tr[:, 1:, :] = treatments[:, :-1, :]
            self.data['current_treatments'] = treatments
        else:
            raise ValueError(self.treatment_mode)
        self.data['prev_treatments'] = prev_tr   # zero row for t = 0, then current_treatments[:, :-1] (:141, :183-185)

        current_covariates = np.empty((R, width - 1, 2))
        current_covariates[..., 0] = cancer_volume[:, :-1]
        current_covariates[..., 1] = patient_types[:, None]
        outputs = cancer_volume[:, 1:, np.newaxis]
        output_means = mean[['cancer_volume']].values.flatten()[0]
        output_stds = std[['cancer_volume']].values.flatten()[0]
        active = (np.arange(outputs.shape[1])[None, :] < seq_len.astype(np.int64)[:, None]).astype(np.float64)

        self.data['current_covariates'] = current_covariates
        self.data['outputs'] = outputs
        self.data['active_entries'] = active[:, :, None]
        self.data['unscaled_outputs'] = outputs * std['cancer_volume'] + mean['cancer_volume']
        self.scaling_params = {'input_means': input_means, 'inputs_stds': input_stds,
                               'output_means': output_means, 'output_stds': output_stds}
        self.data['prev_outputs'] = current_covariates[:, :, :1]
        self.data['static_features'] = current_covariates[:, 0, 1:]
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
        seq_len = self.data['sequence_lengths'].astype(np.int64)
        outputs = self.data['outputs']
        cur = self.data['current_treatments']
        prev = self.data['prev_treatments'][:, 1:, :]
        cov = self.data['current_covariates']
        R, W, _ = outputs.shape
        fact = seq_len - H
        rows = np.arange(R)[:, None]
        k = np.arange(H)[None, :]
        prev_idx = np.mod(fact[:, None] - 1 + k, prev.shape[1])   # python slices with a negative start never occur (sl > H)
        seq = {
            'active_encoder_r': (np.arange(W - H)[None, :] < fact[:, None]).astype(np.float64),
            'prev_treatments': prev[rows, prev_idx, :],
            'current_treatments': cur[rows, fact[:, None] + k, :],
            'current_covariates': np.repeat(cov[np.arange(R), fact - 1][:, None, :], H, axis=1),
            'outputs': outputs[rows, fact[:, None] + k, :],
            'sequence_lengths': np.full(R, float(H)),
            'active_entries': np.ones((R, H, 1)),
        }
        seq['prev_outputs'] = seq['current_covariates'][:, :, :1]
        seq['static_features'] = seq['current_covariates'][:, 0, 1:]
        seq['unscaled_outputs'] = seq['outputs'] * self.scaling_params['output_stds'] + self.scaling_params['output_means']
        seq['patient_types'] = self.data['patient_types']
        seq['patient_ids_all_trajectories'] = self.data['patient_ids_all_trajectories']
        seq['patient_current_t'] = self.data['patient_current_t']
        self.data_original = dict(self.data)     # the reference deep-copies (:470); arrays are shared here
        self.data = seq
        self.processed_sequential = True
        return self.data

    # -- dataset.py:533-552 --------------------------------------------------------------------------
    def process_sequential_multi(self, projection_horizon, include_continuous_treatment=False):
        assert self.processed_sequential
        if not self.processed_autoregressive:
            self.data_processed_seq = self.data
            self.data = dict(self.data_original)   # (:541)
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
