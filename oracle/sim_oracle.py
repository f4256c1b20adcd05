"""TEST INFRASTRUCTURE ONLY.  ctypes front-end of oracle/sim_c/oracle_sim.c.

Each function takes the simulation-parameter dict (reference layout) plus the
pre-drawn random arrays and returns the reference's output dict (same keys / shapes / dtypes as
cancer_simulation.py:356-367, :554-559, :762-769).
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_sim.so")
_lib = None

TUMOUR_DEATH_THRESHOLD = 4 / 3 * np.pi * (13 / 2) ** 3   # cancer_simulation.py:34-35,44

_PARAM_KEYS = ['initial_volumes', 'alpha', 'rho', 'beta', 'beta_c', 'K', 'chemo_sigmoid_intercepts',
               'radio_sigmoid_intercepts', 'chemo_sigmoid_betas', 'radio_sigmoid_betas']


def build(force=False):
    """Compile the C restatement (gcc, a second or two)."""
    src = os.path.join(_HERE, "sim_c", "oracle_sim.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "sim_c")])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_sim_factual.restype = ctypes.c_int
        _lib.oracle_sim_cf_one_step.restype = ctypes.c_int64
        _lib.oracle_sim_cf_treatment_seq.restype = ctypes.c_int64
        _lib.oracle_sim_cf_one_step_windowed.restype = ctypes.c_int64
        _lib.oracle_sim_cf_treatment_seq_windowed.restype = ctypes.c_int64
        _lib.oracle_np_mean.restype = ctypes.c_double
        _lib.oracle_np_sum.restype = ctypes.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def np_mean(a):
    a = _f64(a)
    return lib().oracle_np_mean(_p(a), ctypes.c_int64(a.size))


def sim_factual(params, T, draws, assigned_actions=None, n_threads=1):
    L = lib()
    n = params['initial_volumes'].shape[0]
    pk = [_f64(params[k]) for k in _PARAM_KEYS]
    dr = [_f64(draws[k]) for k in ('noise', 'recovery', 'chemo', 'radio')]
    aa = None if assigned_actions is None else _f64(assigned_actions)
    names = ['cancer_volume', 'chemo_dosage', 'radio_dosage', 'chemo_application', 'radio_application',
             'chemo_probabilities', 'radio_probabilities']
    out = {k: np.empty((n, T)) for k in names}
    out['sequence_lengths'] = np.empty(n)
    out['death_flags'] = np.empty((n, T))
    out['recovery_flags'] = np.empty((n, T))

    def run(lo, hi):
        args = [ctypes.c_int64(hi - lo), ctypes.c_int(T), ctypes.c_int(int(params['window_size'])),
                ctypes.c_int(int(params['lag'])), ctypes.c_double(TUMOUR_DEATH_THRESHOLD)]
        args += [_p(a[lo:hi]) for a in pk]
        args += [_p(a[lo:hi]) for a in dr]
        args += [None if aa is None else _p(aa[lo:hi])]
        args += [_p(out[k][lo:hi]) for k in names]
        args += [_p(out['sequence_lengths'][lo:hi]), _p(out['death_flags'][lo:hi]), _p(out['recovery_flags'][lo:hi])]
        rc = L.oracle_sim_factual(*args)
        assert rc == 0, rc

    if n_threads <= 1 or n < 4 * n_threads:
        run(0, n)
    else:
        bounds = np.linspace(0, n, n_threads + 1).astype(np.int64)
        with ThreadPoolExecutor(n_threads) as ex:
            list(ex.map(lambda b: run(int(b[0]), int(b[1])), zip(bounds[:-1], bounds[1:])))
    out['patient_types'] = params['patient_types']
    assert not np.any(np.isnan(out['cancer_volume'])), 'Cancer volume contains NaN'
    return out


def sim_cf_one_step(params, T, draws, window_rows=None):
    """window_rows (n, T): the output row each patient reads for its treatment window (cancer_simulation.py:471) when
    the n patients are a subset of a larger cohort (row = global patient index); None = the reference verbatim."""
    L = lib()
    n = params['initial_volumes'].shape[0]
    pk = [_f64(params[k]) for k in _PARAM_KEYS]
    pt = _f64(params['patient_types'])
    dr = [_f64(draws[k]) for k in ('noise', 'recovery', 'chemo', 'radio')]
    cap = 4 * n * T
    cv, ca, ra = np.empty((cap, T)), np.empty((cap, T)), np.empty((cap, T))
    sl, pta = np.empty(cap), np.empty(cap)
    tail = [ctypes.c_int64(n), ctypes.c_int(T), ctypes.c_int(int(params['window_size'])),
            ctypes.c_int(int(params['lag'])), ctypes.c_double(TUMOUR_DEATH_THRESHOLD),
            *[_p(a) for a in pk], _p(pt), *[_p(a) for a in dr], _p(cv), _p(ca), _p(ra), _p(sl), _p(pta)]
    if window_rows is None:
        rows = L.oracle_sim_cf_one_step(*tail)
    else:
        wr = _f64(window_rows)
        assert wr.shape == (n, T)
        rows = L.oracle_sim_cf_one_step_windowed(_p(wr), *tail)
    assert rows >= 0, rows
    return {'cancer_volume': cv[:rows], 'chemo_application': ca[:rows], 'radio_application': ra[:rows],
            'sequence_lengths': sl[:rows], 'patient_types': pta[:rows]}


def sim_cf_treatment_seq(params, T, H, draws, window_rows=None):
    """window_rows (n, T+H): see sim_cf_one_step (cancer_simulation.py:671)."""
    L = lib()
    n = params['initial_volumes'].shape[0]
    pk = [_f64(params[k]) for k in _PARAM_KEYS]
    pt = _f64(params['patient_types'])
    dr = [_f64(draws[k]) for k in ('noise', 'recovery', 'chemo', 'radio')]
    cap = 2 * H * n * T
    W = T + H
    cv, ca, ra = np.empty((cap, W)), np.empty((cap, W)), np.empty((cap, W))
    sl, pta, pid, pct = np.empty(cap), np.empty(cap), np.empty(cap), np.empty(cap)
    tail = [ctypes.c_int64(n), ctypes.c_int(T), ctypes.c_int(H),
            ctypes.c_int(int(params['window_size'])), ctypes.c_int(int(params['lag'])),
            ctypes.c_double(TUMOUR_DEATH_THRESHOLD),
            *[_p(a) for a in pk], _p(pt), *[_p(a) for a in dr],
            _p(cv), _p(ca), _p(ra), _p(sl), _p(pta), _p(pid), _p(pct)]
    if window_rows is None:
        rows = L.oracle_sim_cf_treatment_seq(*tail)
    else:
        wr = _f64(window_rows)
        assert wr.shape == (n, W)
        rows = L.oracle_sim_cf_treatment_seq_windowed(_p(wr), *tail)
    assert rows >= 0, rows
    return {'cancer_volume': cv[:rows], 'chemo_application': ca[:rows], 'radio_application': ra[:rows],
            'sequence_lengths': sl[:rows], 'patient_types': pta[:rows],
            'patient_ids_all_trajectories': pid[:rows], 'patient_current_t': pct[:rows]}


def scaling_params(sim):
    """get_scaling_params (:776-796) -> (means dict, stds dict)."""
    L = lib()
    sl = _f64(sim['sequence_lengths'])
    n, T = sim['cancer_volume'].shape
    scratch = np.empty(int(sl.sum()) + 1)
    means, stds = {}, {}
    for k in ['cancer_volume', 'chemo_dosage', 'radio_dosage']:
        m, s = ctypes.c_double(), ctypes.c_double()
        L.oracle_scaling_moments(ctypes.c_int64(n), ctypes.c_int(T), _p(_f64(sim[k])), _p(sl), _p(scratch),
                                 ctypes.byref(m), ctypes.byref(s))
        means[k], stds[k] = m.value, s.value
    means['patient_types'] = np.mean(sim['patient_types'])
    stds['patient_types'] = np.std(sim['patient_types'])
    return means, stds


def theta_gram(sim, static_feature, dt=10.0 / 60):
    """Normal equations of the population fit in long double (C): (G (4,4,4), b (4,4), counts (4,))."""
    L = lib()
    vol = _f64(sim['cancer_volume'])
    n, T = vol.shape
    G, b, cnt = np.zeros((4, 4, 4)), np.zeros((4, 4)), np.zeros(4)
    L.oracle_theta_gram(ctypes.c_int64(n), ctypes.c_int(T), ctypes.c_double(dt), _p(vol),
                        _p(_f64(sim['chemo_application'])), _p(_f64(sim['radio_application'])),
                        _p(_f64(sim['sequence_lengths'])), _p(_f64(static_feature)), _p(G), _p(b), _p(cnt))
    return G, b, cnt


def stlsq_from_gram(G, b, threshold=1e-3, alpha=0.5, max_iter=100):
    """pysindy STLSQ + unbias on normal equations (same loop as sindy_np.stlsq_fit; the ridge step is
    what sklearn's 'cholesky' solver computes: solve(G + alpha I, b))."""
    coefs, sup = np.zeros((4, 4)), np.zeros((4, 4), dtype=bool)
    for a in range(4):
        ind = np.ones(4, dtype=bool)
        prev = np.ones(4, dtype=bool)
        n0 = 4
        c = np.zeros(4)
        if G[a, 0, 0] == 0:
            continue
        for _ in range(max_iter):
            if not ind.any():
                c = np.zeros(4)
                break
            ci = np.linalg.solve(G[a][np.ix_(ind, ind)] + alpha * np.eye(ind.sum()), b[a][ind])
            c = np.zeros(4)
            c[ind] = ci
            big = np.abs(c) >= threshold
            c[~big] = 0
            ind = big
            pattern = c != 0
            same = np.array_equal(pattern, prev)
            prev = pattern
            if ind.sum() == n0 or same:
                break
        if ind.any():
            c = np.zeros(4)
            c[ind] = np.linalg.solve(G[a][np.ix_(ind, ind)], b[a][ind])
        coefs[a], sup[a] = c, ind
    return coefs, sup
