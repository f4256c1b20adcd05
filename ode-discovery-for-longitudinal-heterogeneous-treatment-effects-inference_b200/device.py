"""Device-resident API of the INSITE hot path: thin, typed wrappers over the C ABI.

PyTorch is used for device memory, streams and (in distributed.py) the process group only; every
numeric kernel is hand-written CUDA inside libb200insite.so.  All tensors are float64, C-contiguous
and on the current CUDA device unless stated.  Nothing here falls back to the CPU.
"""
import ctypes
import math

import numpy as np
import torch

from . import _native
from ._native import SimConsts

PARAM_KEYS = ('initial_volumes', 'alpha', 'rho', 'beta', 'beta_c', 'K', 'chemo_sigmoid_intercepts',
              'radio_sigmoid_intercepts', 'chemo_sigmoid_betas', 'radio_sigmoid_betas')
FACTUAL_OUT_KEYS = ('cancer_volume', 'chemo_dosage', 'radio_dosage', 'chemo_application', 'radio_application',
                    'chemo_probabilities', 'radio_probabilities', 'death_flags', 'recovery_flags')

# constants exactly as the reference computes them (cancer_simulation.py:34-44, 231-241)
TUMOUR_CELL_DENSITY = 5.8 * 10 ** 8
TUMOUR_DEATH_THRESHOLD = 4 / 3 * np.pi * (13 / 2) ** 3
SPHERE_COEF = 4 / 3 * np.pi
STANDARD_DT = 10.0 / int(60)          # pkpd/utils.py:48-53
STEPS_FOR_DT = 5                      # pkpd/utils.py:40

STATS_DOUBLES = 68
GRAM_PER_TREATMENT = 15


def sim_consts(window_size=15, lag=0):
    return SimConsts(float(TUMOUR_DEATH_THRESHOLD), float(TUMOUR_CELL_DENSITY), float(SPHERE_COEF), 5.0, 2.0,
                     float(np.exp(-np.log(2) / 1)), int(window_size), int(lag))


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device tensors must be CUDA and contiguous"
    return ctypes.c_void_p(t.data_ptr())


def _ptr_rows(t):
    """Pointer of an (N,T) device tensor whose rows may be pitched (stride(0) >= T, stride(1) == 1)."""
    if t is None:
        return None
    assert t.is_cuda and t.dim() == 2 and (t.shape[1] == 1 or t.stride(1) == 1), "rows must be unit-stride CUDA tensors"
    return ctypes.c_void_p(t.data_ptr())


def row_pitch(*tensors):
    """Common row pitch (elements) of a set of (N,T) tensors; raises if they differ."""
    pitches = {int(t.stride(0)) if t.shape[0] > 1 else int(t.shape[1]) for t in tensors if t is not None}
    if len(pitches) != 1:
        raise ValueError(f"all (N,T) arrays of one call must share their row pitch, got {sorted(pitches)}")
    return pitches.pop()


def alloc_rows(n, T, pitch=None, dtype=torch.float64):
    """(n, T) device tensor; with pitch > T the rows are padded (cudaMallocPitch style).  A pitch that makes a row a
    multiple of 128 bytes (64 for T = 60) keeps the tiled kernels' 16-column boxes on 128-byte lines."""
    pitch = T if pitch is None else int(pitch)
    if pitch == T:
        return torch.empty((n, T), dtype=dtype, device='cuda')
    return torch.empty((n, pitch), dtype=dtype, device='cuda')[:, :T]


def aligned_pitch(T, elem_bytes=8, line=128):
    """Smallest pitch >= T (elements) whose rows are a multiple of `line` bytes."""
    per = line // elem_bytes
    return ((T + per - 1) // per) * per


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("b200_insite needs a CUDA device (B200, sm_100a); there is no CPU path")
    _native.load()


def uniform_param_rows(params):
    """{row index: value} of the parameter rows that hold one scalar for the whole cohort.  The reference's
    generate_params builds K and the four sigmoid rows from scalars (cancer_simulation.py:83-88, :202); a chunked upload need not
    send them.  One pass over the host arrays (a property of the cohort, found once, not per step)."""
    out = {}
    for r, k in enumerate(PARAM_KEYS):
        a = np.asarray(params[k], dtype=np.float64)
        if a.size and a.min() == a.max():
            out[r] = float(a[0])
    return out


def pack_params(params):
    """dict of (N,) arrays (generate_params layout) -> (10,N) float64 numpy block."""
    return np.ascontiguousarray(np.stack([np.asarray(params[k], dtype=np.float64) for k in PARAM_KEYS], axis=0))


def to_device(a, dtype=torch.float64, pinned=False):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype != dtype:
        t = t.to(dtype)
    if pinned:
        t = t.pin_memory()
    return t.cuda(non_blocking=pinned)


_workspaces = {}


def gram_workspace(tag="default"):
    """Per-device statistics workspace (stats live in its first 68 doubles)."""
    lib = _native.load()
    key = (torch.cuda.current_device(), tag)
    if key not in _workspaces:
        nbytes = lib.b200i_gram_workspace_bytes()
        _workspaces[key] = torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device='cuda')
    return _workspaces[key]


def sim_factual(params_dev, noise, recovery, chemo_rvs, radio_rvs, T, consts=None, assigned_actions=None,
                out=None, variant=0, fused_static=None, fd_dt=STANDARD_DT):
    """K1.  Returns (dict of nine (N,T) tensors + 'sequence_lengths' (N,), stats or None).

    fused_static: (N,) un-scaled static feature (patient type); if given the population statistics of
    theta_gram are accumulated inside the simulator kernel and returned as a (68,) tensor view."""
    lib = _native.load()
    n = params_dev.shape[1]
    consts = consts or sim_consts()
    pitch = row_pitch(noise, recovery, chemo_rvs, radio_rvs) if n > 1 else T
    if out is None:
        out = {k: alloc_rows(n, T, pitch) for k in FACTUAL_OUT_KEYS}
        out['sequence_lengths'] = torch.empty((n,), dtype=torch.float64, device='cuda')
    elif n > 1:
        pitch = row_pitch(noise, recovery, chemo_rvs, radio_rvs, *[out[k] for k in FACTUAL_OUT_KEYS])
    ws = gram_workspace() if fused_static is not None else None
    rc = lib.b200i_sim_factual_pitched(n, T, pitch, ctypes.byref(consts), _ptr(params_dev), _ptr_rows(noise),
                                       _ptr_rows(recovery), _ptr_rows(chemo_rvs), _ptr_rows(radio_rvs),
                                       _ptr(assigned_actions), *[_ptr_rows(out[k]) for k in FACTUAL_OUT_KEYS],
                                       _ptr(out['sequence_lengths']), _ptr(fused_static), float(fd_dt), _ptr(ws),
                                       int(variant), _stream())
    _native.check(rc, "b200i_sim_factual")
    return out, (ws[:STATS_DOUBLES] if ws is not None else None)


def sim_factual_side(params_dev, noise, recovery, chemo_rvs, radio_rvs, T, consts=None, out=None, codes=None,
                     patient_moments=None, variant=0):
    """K1 with the side outputs of the lean fit: returns (out dict, codes (N, code_pitch) uint8, moments (6, N))."""
    lib = _native.load()
    n = params_dev.shape[1]
    consts = consts or sim_consts()
    pitch = row_pitch(noise, recovery, chemo_rvs, radio_rvs) if n > 1 else T
    if out is None:
        out = {k: alloc_rows(n, T, pitch) for k in FACTUAL_OUT_KEYS}
        out['sequence_lengths'] = torch.empty((n,), dtype=torch.float64, device='cuda')
    elif n > 1:
        pitch = row_pitch(noise, recovery, chemo_rvs, radio_rvs, *[out[k] for k in FACTUAL_OUT_KEYS])
    if codes is None:
        codes = torch.zeros((n, ((T + 15) // 16) * 16), dtype=torch.uint8, device='cuda')
    if patient_moments is None:
        patient_moments = torch.empty((6, n), dtype=torch.float64, device='cuda')
    rc = lib.b200i_sim_factual_side(n, T, pitch, ctypes.byref(consts), _ptr(params_dev), _ptr_rows(noise),
                                    _ptr_rows(recovery), _ptr_rows(chemo_rvs), _ptr_rows(radio_rvs),
                                    *[_ptr_rows(out[k]) for k in FACTUAL_OUT_KEYS], _ptr(out['sequence_lengths']),
                                    _ptr(codes), int(codes.shape[1]), _ptr(patient_moments), int(variant), _stream())
    _native.check(rc, "b200i_sim_factual_side")
    return out, codes, patient_moments


def philox_draws(n, T, seed, patient_base=0, pitch=None):
    """The device generator's draws as the four (N,T) arrays of the reference contract (noise already x 0.01):
    returns [noise, recovery, chemo, radio].  b200i_sim_factual on these reproduces sim_factual_rng bit for bit."""
    lib = _native.load()
    pitch = T if pitch is None else int(pitch)
    arrs = [alloc_rows(n, T, pitch) for _ in range(4)]
    rc = lib.b200i_philox_draws(n, T, pitch, int(seed), int(patient_base), *[_ptr_rows(a) for a in arrs], _stream())
    _native.check(rc, "b200i_philox_draws")
    return arrs


def sim_factual_rng(params_dev, T, seed, patient_base=0, consts=None, volume=None, codes=None, sequence_lengths=None,
                    patient_moments=None, fused_static=None, fd_dt=STANDARD_DT, pitch=None, moments=True, tag="default",
                    rows=None, variant=0):
    """K1L: simulate_factual with device-generated draws (Philox4x32-10 counted by global patient index).

    Returns (volume (N,T) [row pitch `pitch`], codes (N, code_pitch) uint8 = chemo + 2*radio application,
    sequence_lengths (N,), moments (6,N) or None, stats (68,) or None).  fused_static: (N,) static feature ->
    the population statistics of theta_gram are accumulated inside the kernel (moments are then not written).
    rows=(a, b): simulate only patients [a, b) of the given full-size buffers (chunked pipelines); their generator
    counters are patient_base + a ..., so the result does not depend on the chunking."""
    lib = _native.load()
    n_all = params_dev.shape[1]
    a, b = (0, n_all) if rows is None else (int(rows[0]), int(rows[1]))
    n = b - a
    consts = consts or sim_consts()
    if volume is None:
        volume = alloc_rows(n_all, T, aligned_pitch(T) if pitch is None else pitch)
    vp = row_pitch(volume) if n_all > 1 else T
    if codes is None:
        codes = torch.empty((n_all, ((T + 15) // 16) * 16), dtype=torch.uint8, device='cuda')
    if sequence_lengths is None:
        sequence_lengths = torch.empty((n_all,), dtype=torch.float64, device='cuda')
    ws = gram_workspace(tag) if fused_static is not None else None
    if ws is None and moments and patient_moments is None:
        patient_moments = torch.empty((6, n_all), dtype=torch.float64, device='cuda')
    if ws is not None:
        patient_moments = None
    cp = int(codes.shape[1])

    def off(t, elems, size=8):
        return None if t is None else ctypes.c_void_p(t.data_ptr() + elems * size)
    assert params_dev.is_contiguous() and codes.is_contiguous() and sequence_lengths.is_contiguous()
    rc = lib.b200i_sim_factual_rng(n, T, vp, ctypes.byref(consts), off(params_dev, a), n_all, int(seed),
                                   int(patient_base) + a, off(volume, a * vp), off(codes, a * cp, 1), cp,
                                   off(sequence_lengths, a), off(patient_moments, a), n_all, off(fused_static, a),
                                   float(fd_dt), _ptr(ws), int(variant), _stream())
    _native.check(rc, "b200i_sim_factual_rng")
    return volume, codes, sequence_lengths, patient_moments, (ws[:STATS_DOUBLES] if ws is not None else None)


def chunk_workspaces(chunks):
    """Per-chunk statistics workspaces for upload_simulate_rng (zero-initialised once; the kernels reset them)."""
    lib = _native.load()
    return torch.zeros(int(chunks) * ((lib.b200i_gram_workspace_bytes() + 7) // 8), dtype=torch.float64, device='cuda')


def cohort_scalar_rows(chemo_coeff, radio_coeff):
    """{row index: value} of the parameter rows generate_params fills with one scalar for the whole cohort
    (cancer_simulation.py:83-88, :202): K = calc_volume(30) and the sigmoid intercepts D_MAX / 2 and slopes gamma / D_MAX.
    They are functions of generate_params' ARGUMENTS, so a caller that knows (chemo_coeff, radio_coeff) need not ship
    or scan the five (N,) arrays."""
    d_max = ((TUMOUR_DEATH_THRESHOLD / (4 / 3 * np.pi)) ** (1 / 3)) * 2
    return {5: float(4 / 3 * np.pi * (30 / 2) ** 3), 6: float(d_max / 2.0), 7: float(d_max / 2.0),
            8: float(chemo_coeff / d_max), 9: float(radio_coeff / d_max)}


def upload_simulate_rng(params_host, static_host, params_dev, static_dev, T, seed, patient_base, consts, volume, codes,
                        sequence_lengths, patient_moments, chunks, copy_stream, chunk_ws=None, stats_out=None,
                        fd_dt=STANDARD_DT, uniform=None, derive_beta=False, types_u8_host=None, types_u8_dev=None,
                        pipelined=False):
    """Pinned host parameters -> chunked H2D on copy_stream, each chunk simulated (K1L) on the current stream as soon
    as it has arrived (b200i_upload_simulate_rng).  chunk_ws (chunk_workspaces(chunks)) + stats_out (68,): each
    chunk's share of the population statistics is computed right behind its simulation and summed in chunk order.
    pipelined (reduced input set only): consecutive calls overlap -- this call's copies wait, chunk by chunk, only for
    the previous call's readers of that chunk (b200i_upload_simulate_rng_pipelined)."""
    lib = _native.load()
    n = params_dev.shape[1]
    assert params_host.is_pinned() and params_host.is_contiguous() and tuple(params_host.shape) == (10, n)
    assert static_host is None or (static_host.is_pinned() and static_host.is_contiguous())
    vp = row_pitch(volume) if n > 1 else T
    mask, vals = 0, (ctypes.c_double * 10)()
    for r, v in (uniform or {}).items():     # parameter rows that are one scalar for the whole cohort: not copied
        mask |= 1 << int(r)
        vals[int(r)] = float(v)
    if types_u8_host is not None:
        # reduced parameter set: beta derived from alpha, patient types as bytes (b200i_upload_simulate_rng_reduced)
        assert types_u8_host.is_pinned() and types_u8_host.dtype == torch.uint8 and types_u8_host.numel() == n
        assert types_u8_dev is not None and types_u8_dev.is_cuda and types_u8_dev.numel() == n
        fn = lib.b200i_upload_simulate_rng_pipelined if pipelined else lib.b200i_upload_simulate_rng_reduced
        rc = fn(n, T, vp, ctypes.byref(consts), ctypes.c_void_p(params_host.data_ptr()),
                                                   mask, vals if mask else None, 1 if derive_beta else 0,
                                                   ctypes.c_void_p(types_u8_host.data_ptr()), _ptr(types_u8_dev),
                                                   _ptr(params_dev), _ptr(static_dev), int(seed), int(patient_base),
                                                   _ptr_rows(volume), _ptr(codes), int(codes.shape[1]),
                                                   _ptr(sequence_lengths), _ptr(patient_moments), int(chunks), float(fd_dt),
                                                   _ptr(chunk_ws), _ptr(stats_out),
                                                   ctypes.c_void_p(copy_stream.cuda_stream), _stream())
        _native.check(rc, "b200i_upload_simulate_rng_pipelined" if pipelined else "b200i_upload_simulate_rng_reduced")
        return
    if pipelined:
        raise ValueError("pipelined uploads take the reduced input set (types_u8_host)")
    rc = lib.b200i_upload_simulate_rng(n, T, vp, ctypes.byref(consts), ctypes.c_void_p(params_host.data_ptr()),
                                       mask, vals if mask else None,
                                       None if static_host is None else ctypes.c_void_p(static_host.data_ptr()),
                                       _ptr(params_dev), _ptr(static_dev), int(seed), int(patient_base), _ptr_rows(volume),
                                       _ptr(codes), int(codes.shape[1]), _ptr(sequence_lengths), _ptr(patient_moments),
                                       int(chunks), float(fd_dt), _ptr(chunk_ws), _ptr(stats_out),
                                       ctypes.c_void_p(copy_stream.cuda_stream), _stream())
    _native.check(rc, "b200i_upload_simulate_rng")


def theta_gram_codes(cancer_volume, codes, sequence_lengths, static_feature, patient_moments, fd_dt=STANDARD_DT,
                     tag="default", joint=False):
    """K4 from the simulator's side outputs (see sim_factual_side).  Returns the (68,) packed statistics."""
    lib = _native.load()
    n, T = cancer_volume.shape
    ws = gram_workspace(tag)
    pitch = row_pitch(cancer_volume) if n > 1 else T
    rc = lib.b200i_theta_gram_codes(n, T, pitch, 1 if joint else 0, float(fd_dt), _ptr_rows(cancer_volume), _ptr(codes),
                                    int(codes.shape[1]), _ptr(sequence_lengths), _ptr(static_feature),
                                    ctypes.c_void_p(patient_moments.data_ptr()), int(patient_moments.stride(0)), _ptr(ws),
                                    _stream())
    _native.check(rc, "b200i_theta_gram_codes")
    return ws[:STATS_DOUBLES]


def theta_gram(cancer_volume, chemo_application, radio_application, sequence_lengths, static_feature,
               chemo_dosage=None, radio_dosage=None, fd_dt=STANDARD_DT, tag="default", joint=False):
    """K4.  Returns the (68,) packed statistics (view into the workspace).  joint: the joint model's reduction
    (one trajectory per patient over the outputs, see include/b200i.h)."""
    lib = _native.load()
    n, T = cancer_volume.shape
    ws = gram_workspace(tag)
    pitch = row_pitch(cancer_volume, chemo_application, radio_application, chemo_dosage, radio_dosage) if n > 1 else T
    rc = lib.b200i_theta_gram_mode(n, T, pitch, 1 if joint else 0, float(fd_dt), _ptr_rows(cancer_volume), _ptr_rows(chemo_application),
                                      _ptr_rows(radio_application), _ptr(sequence_lengths), _ptr(static_feature),
                                      _ptr_rows(chemo_dosage), _ptr_rows(radio_dosage), _ptr(ws), _stream())
    _native.check(rc, "b200i_theta_gram")
    return ws[:STATS_DOUBLES]


POLY_TERMS = 15
# PolynomialLibrary(degree=4, interaction_only=False).get_feature_names() over [x0, u0] (sklearn's monomial order)
POLY_FEATURE_LIBRARY_NAMES = ('1', 'x0', 'u0', 'x0^2', 'x0 u0', 'u0^2', 'x0^3', 'x0^2 u0', 'x0 u0^2', 'u0^3',
                              'x0^4', 'x0^3 u0', 'x0^2 u0^2', 'x0 u0^3', 'u0^4')


def poly_tsqr(cancer_volume, chemo_application, radio_application, sequence_lengths, static_feature, fd_dt=STANDARD_DT):
    """Degree-4 library statistics (sindy.py:185-186): the four 16x16 R factors of [Theta | xdot] (tall-skinny QR) and
    the four sample counts, (4*256 + 4,) float64 on the device.  Dense (N,T) rows."""
    lib = _native.load()
    n, T = cancer_volume.shape
    for a in (cancer_volume, chemo_application, radio_application):
        if not a.is_contiguous() or a.shape != cancer_volume.shape:
            raise ValueError("poly_tsqr takes dense (N,T) rows")
    key = (torch.cuda.current_device(), "poly")
    if key not in _workspaces:
        _workspaces[key] = torch.zeros((lib.b200i_poly_workspace_bytes() + 7) // 8, dtype=torch.float64, device='cuda')
    out = torch.empty(4 * 256 + 4, dtype=torch.float64, device='cuda')
    rc = lib.b200i_poly_tsqr(n, T, float(fd_dt), _ptr(cancer_volume), _ptr(chemo_application), _ptr(radio_application),
                             _ptr(sequence_lengths), _ptr(static_feature), _ptr(_workspaces[key]), _ptr(out), _stream())
    _native.check(rc, "b200i_poly_tsqr")
    return out


def poly_stlsq(r_factors, threshold=1e-3, alpha=0.5, max_iter=100, rcond=0.0):
    """pysindy STLSQ + unbias on the R factors -> (coefs (4,15) float64, support (4,15) int32) on the device.
    rcond <= 0: eps * max(samples, 15) (numpy.linalg.lstsq's default cut-off; see include/b200i.h)."""
    lib = _native.load()
    coefs = torch.empty((4, POLY_TERMS), dtype=torch.float64, device='cuda')
    support = torch.empty((4, POLY_TERMS), dtype=torch.int32, device='cuda')
    rc = lib.b200i_poly_stlsq(_ptr(r_factors), float(threshold), float(alpha), int(max_iter), float(rcond), _ptr(coefs),
                              _ptr(support), _stream())
    _native.check(rc, "b200i_poly_stlsq")
    return coefs, support


def poly_rollout(x0, static_feature, codes, coefs, dt=STANDARD_DT, substeps=STEPS_FOR_DT, drop_below=1e-3):
    """K6 for the degree-4 library: codes (R,W) uint8, coefs (4,15) -> (R,W) un-scaled predictions."""
    lib = _native.load()
    rows, W = codes.shape
    out = torch.empty((rows, W), dtype=torch.float64, device='cuda')
    rc = lib.b200i_poly_rollout(rows, W, float(dt), int(substeps), _ptr(x0), _ptr(static_feature), _ptr(codes), _ptr(coefs),
                                float(drop_below), _ptr(out), _stream())
    _native.check(rc, "b200i_poly_rollout")
    return out


def smooth_snippets(cancer_volume, chemo_application, radio_application, sequence_lengths, joint=False):
    """model.use_smoothed_finite_difference pre-pass (sindy.py:196-198): the half-sample two-point Savitzky-Golay mean
    over every fitting trajectory, edges kept (see include/b200i.h).  Dense (N,T) float64 in, new (N,T) tensor out."""
    lib = _native.load()
    n, T = cancer_volume.shape
    for a in (cancer_volume, chemo_application, radio_application):
        if not a.is_contiguous() or a.shape != cancer_volume.shape:
            raise ValueError("smooth_snippets takes dense (N,T) rows")
    out = torch.empty_like(cancer_volume)
    rc = lib.b200i_smooth_snippets(n, T, _ptr(cancer_volume), _ptr(chemo_application), _ptr(radio_application),
                                   _ptr(sequence_lengths), 1 if joint else 0, _ptr(out), _stream())
    _native.check(rc, "b200i_smooth_snippets")
    return out


def theta_gram_dts(cancer_volume, chemo_application, radio_application, sequence_lengths, static_feature, dts,
                   tag="default"):
    """K4 on an irregular time grid: dts (T-1,) or (N,T-1) interval lengths.  Returns the (68,) packed statistics
    (Gram part; the dosage moments are not part of this call)."""
    lib = _native.load()
    n, T = cancer_volume.shape
    ws = gram_workspace(tag)
    dp, dper = _dts_args(dts, n, T - 1)
    rc = lib.b200i_theta_gram_dts(n, T, _ptr(cancer_volume), _ptr(chemo_application), _ptr(radio_application),
                                  _ptr(sequence_lengths), _ptr(static_feature), dp, dper, _ptr(ws), _stream())
    _native.check(rc, "b200i_theta_gram_dts")
    return ws[:STATS_DOUBLES]


def stlsq_population(stats, threshold=1e-3, alpha=0.5, max_iter=100):
    """K5.  stats (68,) device -> (coefs (4,4) float64, support (4,4) int32), both on device."""
    lib = _native.load()
    coefs = torch.empty((4, 4), dtype=torch.float64, device='cuda')
    support = torch.empty((4, 4), dtype=torch.int32, device='cuda')
    rc = lib.b200i_stlsq_population(_ptr(stats), float(threshold), float(alpha), int(max_iter), _ptr(coefs),
                                    _ptr(support), _stream())
    _native.check(rc, "b200i_stlsq_population")
    return coefs, support


def stlsq_joint(stats, threshold=1e-3, alpha=0.5, max_iter=100, drop_below=1e-3):
    """K5j.  stats (68,) from theta_gram(joint=True) -> (coefs (11,), support (11,) int32, per-treatment (4,4)
    coefficients of the thresholded expression for ode_rollout(drop_below=-1))."""
    lib = _native.load()
    coefs = torch.empty((11,), dtype=torch.float64, device='cuda')
    support = torch.empty((11,), dtype=torch.int32, device='cuda')
    c44 = torch.empty((4, 4), dtype=torch.float64, device='cuda')
    rc = lib.b200i_stlsq_joint(_ptr(stats), float(threshold), float(alpha), int(max_iter), float(drop_below),
                               _ptr(coefs), _ptr(support), _ptr(c44), _stream())
    _native.check(rc, "b200i_stlsq_joint")
    return coefs, support, c44


def treatment_codes(chemo_application, radio_application, W):
    lib = _native.load()
    rows, pitch = chemo_application.shape
    codes = torch.empty((rows, W), dtype=torch.uint8, device='cuda')
    rc = lib.b200i_treatment_codes(rows, W, pitch, _ptr(chemo_application), _ptr(radio_application), _ptr(codes),
                                   _stream())
    _native.check(rc, "b200i_treatment_codes")
    return codes


def _dts_args(dts, rows, W):
    """(pointer, per_row flag) of an interval-length array: (W,) shared by all rows or (rows, W) per row."""
    assert dts.is_cuda and dts.is_contiguous() and dts.dtype == torch.float64
    if dts.dim() == 1:
        assert dts.shape[0] == W, f"dts must hold {W} interval lengths, got {tuple(dts.shape)}"
        return _ptr(dts), 0
    assert tuple(dts.shape) == (rows, W), f"dts must be ({rows}, {W}), got {tuple(dts.shape)}"
    return _ptr(dts), 1


def intervals_from_times(t):
    """Time grid(s) t (..., W+1) -> interval lengths (..., W) = diff(t) (what odeint computes first, pkpd/utils.py:86)."""
    return (t[..., 1:] - t[..., :-1]).contiguous()


def ode_rollout(x0, static_feature, codes, coefs, dt=STANDARD_DT, substeps=STEPS_FOR_DT, drop_below=1e-3, out=None,
                fp32=False, dts=None):
    """K6.  codes (R,W) uint8; coefs (4,4) or (R,4,4).  Returns (R,W) un-scaled predictions.
    fp32: integrate in float32 (inputs / outputs stay float64).
    dts: interval lengths (W,) or (R,W) for an irregular time grid (then `dt` is ignored)."""
    lib = _native.load()
    rows, W = codes.shape
    per_row = 1 if coefs.dim() == 3 else 0
    if out is None:
        out = torch.empty((rows, W), dtype=torch.float64, device='cuda')
    if dts is not None:
        dp, dper = _dts_args(dts, rows, W)
        rc = lib.b200i_ode_rollout_dts(rows, W, int(substeps), _ptr(x0), _ptr(static_feature), _ptr(codes), _ptr(coefs),
                                       per_row, float(drop_below), dp, dper, 1 if fp32 else 0, _ptr(out), _stream())
        _native.check(rc, "b200i_ode_rollout_dts")
        return out
    fn = lib.b200i_ode_rollout_f32 if fp32 else lib.b200i_ode_rollout
    rc = fn(rows, W, float(dt), int(substeps), _ptr(x0), _ptr(static_feature), _ptr(codes),
            _ptr(coefs), per_row, float(drop_below), _ptr(out), _stream())
    _native.check(rc, "b200i_ode_rollout")
    return out


def stlsq_batched(x, codes, fit_len, static_feature, prior, lam, threshold=1e-3, support_tol=1e-3, max_iter=10,
                  fd_dt=STANDARD_DT, dts=None, estimator='ridge_prior'):
    """K5b.  x (R,W) float64 (or float32: FP32 storage, FP64 arithmetic), codes (R,W) uint8, fit_len (R,) int32 ->
    per-row coefficients (R,4,4).  dts: interval lengths (W,) or (R,W) for an irregular time grid.
    estimator: 'ridge_prior' (north star) or 'lsq_initial_mask' (the ridge / threshold loop of the reference's dormant
    LSQIntialMask: prior = initial support only, lam = its alpha, pass support_tol=1e-14; see include/b200i.h)."""
    lib = _native.load()
    rows, W = x.shape
    est = {'ridge_prior': 0, 'lsq_initial_mask': 1}[estimator]
    out = torch.empty((rows, 4, 4), dtype=torch.float64, device='cuda')
    if dts is not None or x.dtype == torch.float32 or est != 0:
        dp, dper = (None, 0) if dts is None else _dts_args(dts, rows, W)
        f32 = x.dtype == torch.float32
        assert x.is_contiguous()
        rc = lib.b200i_stlsq_batched_dts(rows, W, None if f32 else _ptr(x), _ptr(x) if f32 else None, _ptr(codes),
                                         _ptr(fit_len), _ptr(static_feature), _ptr(prior), float(support_tol), float(lam),
                                         float(threshold), int(max_iter), float(fd_dt), dp, dper, est, _ptr(out), _stream())
        _native.check(rc, "b200i_stlsq_batched_dts")
        return out
    rc = lib.b200i_stlsq_batched(rows, W, float(fd_dt), _ptr(x), _ptr(codes), _ptr(fit_len), _ptr(static_feature),
                                 _ptr(prior), float(support_tol), float(lam), float(threshold), int(max_iter),
                                 _ptr(out), _stream())
    _native.check(rc, "b200i_stlsq_batched")
    return out


LINE_SEARCH = {'robust': 0, 'jax': 1}      # B200I_LS_ROBUST / B200I_LS_JAX
JAX_GTOL = 1e-5                            # minimize_bfgs' default: jax.scipy.optimize.minimize does not forward `tol`


def _bfgs_options(line_search, gtol, max_iter, n_coefs):
    """Defaults of the two line-search flavours: 'jax' = what the reference's minimize(..., method='BFGS', tol=1e-12)
    really runs (gtol 1e-5, maxiter 200 * n); 'robust' = tight tolerance, best-point acceptance."""
    mode = LINE_SEARCH[line_search]
    if gtol is None:
        gtol = JAX_GTOL if mode == 1 else 1e-12
    if max_iter is None:
        max_iter = 200 * n_coefs if mode == 1 else 200
    return mode, float(gtol), int(max_iter)


def insite_bfgs(x, codes, sequence_lengths, projection_horizon, static_feature, theta0, lam, gtol=None,
                max_iter=None, dt=STANDARD_DT, substeps=STEPS_FOR_DT, joint=False, dts=None, line_search='jax'):
    """K7.  Returns (coefs (R,4,4) [joint: (R,11)], status (R,) int32, fval (R,2)).
    line_search: 'jax' (the reference's optimiser, restated; status 3 rows are the ones sindy.py:628-631 replaces by
    theta0) or 'robust'; gtol / max_iter default per flavour (see _bfgs_options).
    dts: interval lengths (W,) or (R,W) for an irregular time grid (per-treatment models)."""
    lib = _native.load()
    rows, W = x.shape
    mode, gtol, max_iter = _bfgs_options(line_search, gtol, max_iter, 11 if joint else 16)
    if dts is not None:
        assert not joint, "irregular grids are implemented for the per-treatment models"
        dp, dper = _dts_args(dts, rows, W)
        coefs = torch.empty((rows, 4, 4), dtype=torch.float64, device='cuda')
        status = torch.empty((rows,), dtype=torch.int32, device='cuda')
        fval = torch.empty((rows, 2), dtype=torch.float64, device='cuda')
        rc = lib.b200i_insite_bfgs_dts(rows, W, int(substeps), _ptr(x), _ptr(codes), _ptr(sequence_lengths),
                                       int(projection_horizon), _ptr(static_feature), _ptr(theta0), float(lam), float(gtol),
                                       int(max_iter), mode, dp, dper, _ptr(coefs), _ptr(status), _ptr(fval), _stream())
        _native.check(rc, "b200i_insite_bfgs_dts")
        return coefs, status, fval
    if joint:
        coefs = torch.empty((rows, 11), dtype=torch.float64, device='cuda')
        status = torch.empty((rows,), dtype=torch.int32, device='cuda')
        fval = torch.empty((rows, 2), dtype=torch.float64, device='cuda')
        rc = lib.b200i_insite_bfgs_joint(rows, W, float(dt), int(substeps), _ptr(x), _ptr(codes), _ptr(sequence_lengths),
                                         int(projection_horizon), _ptr(static_feature), _ptr(theta0), float(lam),
                                         float(gtol), int(max_iter), mode, _ptr(coefs), _ptr(status), _ptr(fval), _stream())
        _native.check(rc, "b200i_insite_bfgs_joint")
        return coefs, status, fval
    coefs = torch.empty((rows, 4, 4), dtype=torch.float64, device='cuda')
    status = torch.empty((rows,), dtype=torch.int32, device='cuda')
    fval = torch.empty((rows, 2), dtype=torch.float64, device='cuda')
    rc = lib.b200i_insite_bfgs(rows, W, float(dt), int(substeps), _ptr(x), _ptr(codes), _ptr(sequence_lengths),
                               int(projection_horizon), _ptr(static_feature), _ptr(theta0), float(lam), float(gtol),
                               int(max_iter), mode, _ptr(coefs), _ptr(status), _ptr(fval), _stream())
    _native.check(rc, "b200i_insite_bfgs")
    return coefs, status, fval


_mse_ws = {}


def masked_se(pred, target, active_len):
    """(3W+2,) sums: se per column, active count per column, last-entry se per column, total last se, #rows."""
    lib = _native.load()
    rows, W = pred.shape
    dev = torch.cuda.current_device()
    if dev not in _mse_ws:
        _mse_ws[dev] = torch.zeros((lib.b200i_masked_se_workspace_bytes() + 7) // 8, dtype=torch.float64, device='cuda')
    sums = torch.empty(3 * W + 2, dtype=torch.float64, device='cuda')
    rc = lib.b200i_masked_se(rows, W, _ptr(pred), _ptr(target), _ptr(active_len), _ptr(sums), _ptr(_mse_ws[dev]),
                             _stream())
    _native.check(rc, "b200i_masked_se")
    return sums


def unpack_stats(stats):
    """(68,) host array -> dict with per-treatment G (4,4,4), b (4,4), counts (4,), moments."""
    s = np.asarray(stats, dtype=np.float64)
    G = np.zeros((4, 4, 4)); b = np.zeros((4, 4)); cnt = np.zeros(4)
    for a in range(4):
        g = s[a * GRAM_PER_TREATMENT:(a + 1) * GRAM_PER_TREATMENT]
        k = 0
        for i in range(4):
            for j in range(i, 4):
                G[a, i, j] = G[a, j, i] = g[k]
                k += 1
        b[a] = g[10:14]
        cnt[a] = g[14]
    m = s[4 * GRAM_PER_TREATMENT:]
    return {'G': G, 'b': b, 'count': cnt,
            'sum_v': m[0], 'sum_vv': m[1], 'sum_c': m[2], 'sum_cc': m[3], 'sum_d': m[4], 'sum_dd': m[5],
            'active': m[6], 'patients': m[7]}


def moments_to_scaling(stats):
    """mean / std (ddof=0) of cancer_volume, chemo_dosage, radio_dosage over active entries
    (get_scaling_params, cancer_simulation.py:776-796) from the packed statistics."""
    u = unpack_stats(stats)
    n = u['active']
    out = {}
    for name, s1, s2 in (('cancer_volume', 'sum_v', 'sum_vv'), ('chemo_dosage', 'sum_c', 'sum_cc'),
                         ('radio_dosage', 'sum_d', 'sum_dd')):
        mean = u[s1] / n
        var = max(u[s2] / n - mean * mean, 0.0)
        out[name] = (mean, math.sqrt(var))
    return out


# ---- evaluation on compact counterfactual cohorts (csrc/cf_eval.cu) ---------------------------------------------------
def cf_eval_one_step(factual, codes, cf, n_steps, static_feature, coefs, drop_below=1e-3, dt=STANDARD_DT,
                     substeps=STEPS_FOR_DT):
    """K8.  Compact one-step cohort (factual (n,T), codes (n,T) uint8, cf (n,T-1,4), n_steps (n,) int32) scored against
    the ODE with coefficients (4,4) [whole cohort] or (n,T-1,4,4) [per (patient, t)].  Returns the (3(T-1)+2,) sums of
    masked_se over the reference's dense rows: se per column, active rows per column, last-entry se per column, total
    last-entry se, rows."""
    lib = _native.load()
    n, T = factual.shape
    per_step = 1 if coefs.dim() == 4 else 0
    assert per_step == 0 or tuple(coefs.shape) == (n, T - 1, 4, 4)
    sums = torch.empty(3 * (T - 1) + 2, dtype=torch.float64, device='cuda')
    rc = lib.b200i_cf_eval_one_step(n, T, float(dt), int(substeps), _ptr(factual), _ptr(codes), _ptr(cf), _ptr(n_steps),
                                    _ptr(static_feature), _ptr(coefs), per_step, float(drop_below), _ptr(sums), _stream())
    _native.check(rc, "b200i_cf_eval_one_step")
    return sums


def cf_eval_treatment_seq(factual, codes, cf, valid, n_steps, static_feature, coefs, drop_below=1e-3, dt=STANDARD_DT,
                          substeps=STEPS_FOR_DT):
    """K9.  Compact treatment-sequence cohort (cf (n,T-1,2H,H), valid (n,T-1) bit masks) -> (2H,) sums: squared error per
    projection step, scored rows per projection step."""
    lib = _native.load()
    n, T = factual.shape
    H = cf.shape[3]
    per_step = 1 if coefs.dim() == 4 else 0
    assert per_step == 0 or tuple(coefs.shape) == (n, T - 1, 4, 4)
    sums = torch.empty(2 * H, dtype=torch.float64, device='cuda')
    rc = lib.b200i_cf_eval_treatment_seq(n, T, H, float(dt), int(substeps), _ptr(factual), _ptr(codes), _ptr(cf), _ptr(valid),
                                         _ptr(n_steps), _ptr(static_feature), _ptr(coefs), per_step, float(drop_below),
                                         _ptr(sums), _stream())
    _native.check(rc, "b200i_cf_eval_treatment_seq")
    return sums


def insite_bfgs_prefix(factual, codes, n_steps, static_feature, theta0, lam, fit_offset, gtol=None, max_iter=None,
                       dt=STANDARD_DT, substeps=STEPS_FOR_DT, line_search='jax'):
    """K7 per (patient, t) of a compact cohort: fit window = the first t + fit_offset transitions of the factual
    trajectory.  Returns (coefs (n,T-1,4,4), status (n,T-1) int32, fval (n,T-1,2))."""
    lib = _native.load()
    n, T = factual.shape
    mode, gtol, max_iter = _bfgs_options(line_search, gtol, max_iter, 16)
    coefs = torch.empty((n, T - 1, 4, 4), dtype=torch.float64, device='cuda')
    status = torch.empty((n, T - 1), dtype=torch.int32, device='cuda')
    fval = torch.empty((n, T - 1, 2), dtype=torch.float64, device='cuda')
    rc = lib.b200i_insite_bfgs_prefix(n, T, int(fit_offset), float(dt), int(substeps), _ptr(factual), _ptr(codes),
                                      _ptr(n_steps), _ptr(static_feature), _ptr(theta0), float(lam), float(gtol),
                                      int(max_iter), mode, _ptr(coefs), _ptr(status), _ptr(fval), _stream())
    _native.check(rc, "b200i_insite_bfgs_prefix")
    return coefs, status, fval


def stlsq_prefix(factual, codes, n_steps, static_feature, prior, lam, fit_offset, threshold=1e-3, support_tol=1e-3,
                 max_iter=10, fd_dt=STANDARD_DT):
    """K5b per (patient, t) of a compact cohort (running sums, one pass per patient) -> coefs (n,T-1,4,4)."""
    lib = _native.load()
    n, T = factual.shape
    coefs = torch.empty((n, T - 1, 4, 4), dtype=torch.float64, device='cuda')
    rc = lib.b200i_stlsq_prefix(n, T, int(fit_offset), float(fd_dt), _ptr(factual), _ptr(codes), _ptr(n_steps),
                                _ptr(static_feature), _ptr(prior), float(support_tol), float(lam), float(threshold),
                                int(max_iter), _ptr(coefs), _stream())
    _native.check(rc, "b200i_stlsq_prefix")
    return coefs
