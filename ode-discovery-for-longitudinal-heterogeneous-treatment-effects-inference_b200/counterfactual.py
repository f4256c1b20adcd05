"""Counterfactual cohorts: device-resident compact representation (K2 / K3) and the conversion to the
reference's dense row layout for run.py-scale cohorts.

Reference: simulate_counterfactual_1_step (cancer_simulation.py:378-563) and
simulate_counterfactuals_treatment_seq (:566-773).  See include/b200i.h for the compact layout and the
cross-row window semantics (:471, :671) that make the cohort a dependency wavefront.
"""
import ctypes

import numpy as np
import torch

from . import _native
from . import device as dev
from ._native import CfSource


class CompactCohort:
    """Per-patient counterfactual cohort on the device."""

    def __init__(self, kind, n, T, H, factual, codes, cf, valid, n_steps, n_rows, row_offsets, total_rows, levels,
                 patient_types=None):
        self.kind, self.n, self.T, self.H = kind, n, T, H
        self.factual, self.codes, self.cf, self.valid = factual, codes, cf, valid
        self.n_steps, self.n_rows, self.row_offsets = n_steps, n_rows, row_offsets
        self.total_rows, self.levels = total_rows, levels
        self.patient_types = patient_types

    def as_source(self):
        return CfSource(self.n, self.factual.data_ptr(), self.codes.data_ptr(), self.cf.data_ptr(),
                        self.valid.data_ptr() if self.valid is not None else None, self.row_offsets.data_ptr())


def sim_cf_one_step(params_dev, noise, recovery, chemo_rvs, radio_rvs, T, consts=None, global_base=0, source=None):
    lib = _native.load()
    n = params_dev.shape[1]
    consts = consts or dev.sim_consts()
    F = torch.empty((n, T), dtype=torch.float64, device='cuda')
    codes = torch.empty((n, T), dtype=torch.uint8, device='cuda')
    cf = torch.empty((n, T - 1, 4), dtype=torch.float64, device='cuda')
    n_steps = torch.empty((n,), dtype=torch.int32, device='cuda')
    n_rows = torch.empty((n,), dtype=torch.int32, device='cuda')
    off = torch.empty((n + 1,), dtype=torch.int64, device='cuda')
    total, levels = ctypes.c_int64(0), ctypes.c_int32(0)
    src = source.as_source() if source is not None else None
    rc = lib.b200i_sim_cf_one_step(n, T, ctypes.byref(consts), dev._ptr(params_dev), dev._ptr(noise),
                                   dev._ptr(recovery), dev._ptr(chemo_rvs), dev._ptr(radio_rvs), int(global_base),
                                   ctypes.byref(src) if src is not None else None,
                                   dev._ptr(F), dev._ptr(codes), dev._ptr(cf), dev._ptr(n_steps), dev._ptr(n_rows),
                                   dev._ptr(off), ctypes.byref(total), ctypes.byref(levels), dev._stream())
    _native.check(rc, "b200i_sim_cf_one_step")
    return CompactCohort('one_step', n, T, 0, F, codes, cf, None, n_steps, n_rows, off, total.value, levels.value)


def sim_cf_treatment_seq(params_dev, noise, recovery, chemo_rvs, radio_rvs, T, H, consts=None, global_base=0,
                         source=None):
    lib = _native.load()
    n = params_dev.shape[1]
    consts = consts or dev.sim_consts()
    F = torch.empty((n, T), dtype=torch.float64, device='cuda')
    codes = torch.empty((n, T), dtype=torch.uint8, device='cuda')
    cf = torch.empty((n, T - 1, 2 * H, H), dtype=torch.float64, device='cuda')
    valid = torch.empty((n, T - 1), dtype=torch.int16, device='cuda')   # bit mask, read as uint16
    n_steps = torch.empty((n,), dtype=torch.int32, device='cuda')
    n_rows = torch.empty((n,), dtype=torch.int32, device='cuda')
    off = torch.empty((n + 1,), dtype=torch.int64, device='cuda')
    total, levels = ctypes.c_int64(0), ctypes.c_int32(0)
    src = source.as_source() if source is not None else None
    rc = lib.b200i_sim_cf_treatment_seq(n, T, H, ctypes.byref(consts), dev._ptr(params_dev), dev._ptr(noise),
                                        dev._ptr(recovery), dev._ptr(chemo_rvs), dev._ptr(radio_rvs),
                                        int(global_base), ctypes.byref(src) if src is not None else None,
                                        dev._ptr(F), dev._ptr(codes), dev._ptr(cf), dev._ptr(valid),
                                        dev._ptr(n_steps), dev._ptr(n_rows), dev._ptr(off), ctypes.byref(total),
                                        ctypes.byref(levels), dev._stream())
    _native.check(rc, "b200i_sim_cf_treatment_seq")
    return CompactCohort('treatment_seq', n, T, H, F, codes, cf, valid, n_steps, n_rows, off, total.value,
                         levels.value)


def generated_draws(n, T, H, seed, patient_base=0):
    """Throughput mode: the per-patient draws of the counterfactual generators (0.01 * randn(T + H), rand(T) x 3;
    cancer_simulation.py:440-453, :640-653) from the device generator instead of numpy's sequential stream --
    Philox4x32-10 counted by (global patient index, column pair, stream), csrc/philox.cuh, so the draws of a patient
    do not depend on the shard or on H.  Returns (noise (n, T+H), recovery, chemo, radio (n, T)) on the device."""
    W = T + H
    Wg = W + (W & 1)                                   # the generator works in column pairs
    noise, rec, chemo, radio = dev.philox_draws(n, Wg, seed, patient_base=patient_base)
    cut = lambda a, w: a if a.shape[1] == w else a[:, :w].contiguous()
    return cut(noise, W), cut(rec, T), cut(chemo, T), cut(radio, T)


def expand(cohort, patient_types_dev, row_begin=0, row_end=None):
    """Dense reference rows [row_begin,row_end) as a dict of device tensors (reference key names)."""
    lib = _native.load()
    row_end = cohort.total_rows if row_end is None else row_end
    R = row_end - row_begin
    W = cohort.T + cohort.H
    f64 = dict(dtype=torch.float64, device='cuda')
    out = {'cancer_volume': torch.empty((R, W), **f64), 'chemo_application': torch.empty((R, W), **f64),
           'radio_application': torch.empty((R, W), **f64), 'sequence_lengths': torch.empty((R,), **f64),
           'patient_types': torch.empty((R,), **f64)}
    if cohort.kind == 'one_step':
        rc = lib.b200i_expand_cf_one_step(cohort.n, cohort.T, dev._ptr(cohort.factual), dev._ptr(cohort.codes),
                                          dev._ptr(cohort.cf), dev._ptr(cohort.row_offsets),
                                          dev._ptr(patient_types_dev), row_begin, row_end,
                                          dev._ptr(out['cancer_volume']), dev._ptr(out['chemo_application']),
                                          dev._ptr(out['radio_application']), dev._ptr(out['sequence_lengths']),
                                          dev._ptr(out['patient_types']), dev._stream())
        _native.check(rc, "b200i_expand_cf_one_step")
    else:
        out['patient_ids_all_trajectories'] = torch.empty((R,), **f64)
        out['patient_current_t'] = torch.empty((R,), **f64)
        rc = lib.b200i_expand_cf_treatment_seq(cohort.n, cohort.T, cohort.H, dev._ptr(cohort.factual),
                                               dev._ptr(cohort.codes), dev._ptr(cohort.cf), dev._ptr(cohort.valid),
                                               dev._ptr(cohort.row_offsets), dev._ptr(patient_types_dev),
                                               row_begin, row_end, dev._ptr(out['cancer_volume']),
                                               dev._ptr(out['chemo_application']), dev._ptr(out['radio_application']),
                                               dev._ptr(out['sequence_lengths']), dev._ptr(out['patient_types']),
                                               dev._ptr(out['patient_ids_all_trajectories']),
                                               dev._ptr(out['patient_current_t']), dev._stream())
        _native.check(rc, "b200i_expand_cf_treatment_seq")
    return out


def _upload(simulation_params, draws):
    params_dev = dev.to_device(dev.pack_params(simulation_params))
    noise, rec, chemo, radio = (dev.to_device(a) for a in draws)
    ptypes = dev.to_device(np.asarray(simulation_params['patient_types'], dtype=np.float64))
    consts = dev.sim_consts(simulation_params['window_size'], simulation_params['lag'])
    return params_dev, noise, rec, chemo, radio, ptypes, consts


def _row_meta(cohort, patient_types):
    """The small per-row arrays of the reference dictionaries from the per-patient arrays (no dense expansion):
    sequence_lengths, patient_types [, patient_ids_all_trajectories, patient_current_t], all float64 as the reference's
    np.zeros buffers (:423-430, :621-630)."""
    T, H = cohort.T, cohort.H
    ns = cohort.n_steps.cpu().numpy().astype(np.int64)
    n = ns.shape[0]
    ptypes = np.asarray(patient_types, dtype=np.float64)
    if cohort.kind == 'one_step':
        per_step = np.where(np.arange(T - 1)[None, :] < ns[:, None], 4, 0)
    else:
        v = cohort.valid.cpu().numpy().view(np.uint16).astype(np.uint32)
        pop = np.zeros_like(v)
        for b in range(2 * H):
            pop += (v >> b) & 1
        per_step = np.where(np.arange(T - 1)[None, :] < ns[:, None], pop, 0).astype(np.int64)
    counts = per_step.reshape(-1)
    pid = np.repeat(np.repeat(np.arange(n), T - 1), counts)
    cur_t = np.repeat(np.tile(np.arange(T - 1), n), counts)
    meta = {'sequence_lengths': (cur_t + 1 + H).astype(np.float64), 'patient_types': ptypes[pid]}
    if cohort.kind != 'one_step':
        meta['patient_ids_all_trajectories'] = pid.astype(np.float64)
        meta['patient_current_t'] = cur_t.astype(np.float64)
    assert meta['sequence_lengths'].shape[0] == cohort.total_rows
    return meta


def _lazy_dense(cohort, ptypes_dev, patient_types):
    """Dictionary with the reference's keys (cancer_simulation.py:554-559 / :762-769).  The three (R, width) arrays -- 227
    resp. 562 rows per test patient -- stay on the device in compact form until somebody reads them: the first access
    of any of them expands the cohort and copies the dense rows to the host once.  The compact cohort itself travels
    as attrs['compact'] = (cohort, static feature on the device): SINDY evaluates it without dense rows."""
    from .lazydict import LazyDict
    big = ('cancer_volume', 'chemo_application', 'radio_application')

    def build():
        dense = expand(cohort, ptypes_dev)
        torch.cuda.current_stream().synchronize()
        return {k: dense[k].cpu().numpy() for k in big}
    out = LazyDict()
    out.set_lazy_group(big, build)
    for k, v in _row_meta(cohort, patient_types).items():
        out[k] = v
    out.attrs['compact'] = (cohort, ptypes_dev)
    return out


def one_step_dense(simulation_params, seq_length, draws, lazy=False):
    """numpy-in / numpy-out body of simulate_counterfactual_1_step (dict keys of :554-559).
    lazy: return the dictionary with the dense arrays pending (see _lazy_dense)."""
    dev.require_cuda()
    params_dev, noise, rec, chemo, radio, ptypes, consts = _upload(simulation_params, draws)
    cohort = sim_cf_one_step(params_dev, noise, rec, chemo, radio, seq_length, consts)
    if lazy:
        return _lazy_dense(cohort, ptypes, simulation_params['patient_types'])
    dense = expand(cohort, ptypes)
    torch.cuda.current_stream().synchronize()
    return {k: dense[k].cpu().numpy() for k in
            ('cancer_volume', 'chemo_application', 'radio_application', 'sequence_lengths', 'patient_types')}


def treatment_seq_dense(simulation_params, seq_length, projection_horizon, draws, lazy=False):
    """numpy-in / numpy-out body of simulate_counterfactuals_treatment_seq (dict keys of :762-769)."""
    dev.require_cuda()
    params_dev, noise, rec, chemo, radio, ptypes, consts = _upload(simulation_params, draws)
    cohort = sim_cf_treatment_seq(params_dev, noise, rec, chemo, radio, seq_length, projection_horizon, consts)
    if lazy:
        return _lazy_dense(cohort, ptypes, simulation_params['patient_types'])
    dense = expand(cohort, ptypes)
    torch.cuda.current_stream().synchronize()
    return {k: dense[k].cpu().numpy() for k in
            ('cancer_volume', 'chemo_application', 'radio_application', 'sequence_lengths', 'patient_types',
             'patient_ids_all_trajectories', 'patient_current_t')}


# ------------------------------------------------------------------------------------------------------------------
# multi-GPU: contiguous patient shards of one counterfactual cohort (the A3/A4 cross-row exception)
# ------------------------------------------------------------------------------------------------------------------
def rows_per_patient_floor(kind, H=5):
    """Rows a patient emits at the very least per executed step (one-step: 4 options; sequences: 2H minus NaN drops,
    in practice 2H) -- only used to size the first guess of the source prefix."""
    return 4 if kind == 'one_step' else 2 * H


def exchange_row_bases(total_rows_local):
    """Exclusive scan of the per-rank row totals over the process group: (row index of this rank's first row, total
    rows of the cohort).  This is the only exchange the sharded generators need (world_size int64 values); without a
    process group it is (0, total_rows_local)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return 0, int(total_rows_local)
    world, rank = dist.get_world_size(), dist.get_rank()
    dev_ = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    mine = torch.tensor([int(total_rows_local)], dtype=torch.int64, device=dev_)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    totals = [int(p.item()) for p in parts]
    return sum(totals[:rank]), sum(totals)


def sim_cf_shard(kind, T, H, n_total, global_base, shard_inputs, prefix_inputs, consts=None, row_base=None):
    """One rank's shard [global_base, global_base + n_shard) of an n_total-patient counterfactual cohort.

    Patient i's treatment probabilities read OUTPUT ROW i of the whole cohort (cancer_simulation.py:471 / :671), a row
    emitted by one of the first ~n_total/227 (one-step) or ~n_total/562 (sequences) patients.  Every rank therefore
    simulates that global source prefix itself (prefix_inputs: device params (10,P) + the four draw arrays of patients
    0..P-1; with device-generated draws that is a regeneration, not a transfer), level by level, and then its own
    shard in ONE launch against the prefix (b200i_cf_source): no collective on the data path.  The reference row
    indices of the shard's rows need the row totals of the ranks before it: exchange_row_bases (one all-gather of an
    int64 per rank), or row_base when the caller already knows it.

    shard_inputs / prefix_inputs: (params_dev, noise, recovery, chemo_rvs, radio_rvs).
    Returns (shard cohort with GLOBAL row_offsets, prefix cohort, total rows of the whole cohort or None)."""
    run = sim_cf_one_step if kind == 'one_step' else sim_cf_treatment_seq
    extra = () if kind == 'one_step' else (H,)
    src = run(*prefix_inputs, T, *extra, consts=consts)
    covered = int(src.row_offsets[-1].item())
    if covered < n_total and src.n < n_total:
        raise ValueError(f"source prefix of {src.n} patients emits {covered} rows but the cohort reads rows up to "
                         f"{n_total}: simulate a longer prefix")
    shard = run(*shard_inputs, T, *extra, consts=consts, global_base=int(global_base), source=src)
    if row_base is None:
        row_base, total = exchange_row_bases(shard.total_rows)
    else:
        total = None
    shard.row_offsets += int(row_base)
    shard.row_base = int(row_base)
    return shard, src, total
