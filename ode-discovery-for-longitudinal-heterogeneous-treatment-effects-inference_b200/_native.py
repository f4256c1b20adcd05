"""ctypes binding of libb200insite.so (C ABI in include/b200i.h).

There is no CPU fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libb200insite.so")

c_i32, c_i64, c_f64, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p


class SimConsts(ctypes.Structure):
    """b200i_sim_consts"""
    _fields_ = [("death_threshold", c_f64), ("cell_density", c_f64), ("sphere_coef", c_f64),
                ("chemo_amt", c_f64), ("radio_amt", c_f64), ("drug_decay", c_f64),
                ("window_size", c_i32), ("lag", c_i32)]


class CfSource(ctypes.Structure):
    """b200i_cf_source"""
    _fields_ = [("n", c_i64), ("factual", c_vp), ("codes", c_vp), ("cf", c_vp), ("valid", c_vp),
                ("row_offsets", c_vp)]


# name -> (restype, argtypes); every symbol declared in include/b200i.h
SIGNATURES = {
    "b200i_last_error": (ctypes.c_char_p, []),
    "b200i_version": (ctypes.c_int, []),
    "b200i_device_sms": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "b200i_sim_factual": (ctypes.c_int, [c_i64, c_i32, ctypes.POINTER(SimConsts), c_vp] + [c_vp] * 5 + [c_vp] * 10 +
                          [c_vp, c_f64, c_vp, c_i32, c_vp]),
    "b200i_sim_factual_pitched": (ctypes.c_int, [c_i64, c_i32, c_i64, ctypes.POINTER(SimConsts), c_vp] + [c_vp] * 5 +
                                  [c_vp] * 10 + [c_vp, c_f64, c_vp, c_i32, c_vp]),
    "b200i_theta_gram_pitched": (ctypes.c_int, [c_i64, c_i32, c_i64, c_f64] + [c_vp] * 9),
    "b200i_theta_gram_mode": (ctypes.c_int, [c_i64, c_i32, c_i64, c_i32, c_f64] + [c_vp] * 9),
    "b200i_stlsq_joint": (ctypes.c_int, [c_vp, c_f64, c_f64, c_i32, c_f64, c_vp, c_vp, c_vp, c_vp]),
    "b200i_sim_factual_side": (ctypes.c_int, [c_i64, c_i32, c_i64, ctypes.POINTER(SimConsts), c_vp] + [c_vp] * 4 +
                               [c_vp] * 10 + [c_vp, c_i64, c_vp, c_i32, c_vp]),
    "b200i_theta_gram_codes": (ctypes.c_int, [c_i64, c_i32, c_i64, c_i32, c_f64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp,
                                              c_i64, c_vp, c_vp]),
    "b200i_philox_draws": (ctypes.c_int, [c_i64, c_i32, c_i64, ctypes.c_uint64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "b200i_sim_factual_rng": (ctypes.c_int, [c_i64, c_i32, c_i64, ctypes.POINTER(SimConsts), c_vp, c_i64, ctypes.c_uint64,
                                             c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_f64, c_vp, c_i32, c_vp]),
    "b200i_upload_simulate_rng": (ctypes.c_int, [c_i64, c_i32, c_i64, ctypes.POINTER(SimConsts), c_vp, ctypes.c_uint32,
                                                 ctypes.POINTER(c_f64), c_vp, c_vp, c_vp,
                                                 ctypes.c_uint64, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_i32, c_f64, c_vp, c_vp,
                                                 c_vp, c_vp]),
    "b200i_upload_simulate_rng_reduced": (ctypes.c_int, [c_i64, c_i32, c_i64, ctypes.POINTER(SimConsts), c_vp, ctypes.c_uint32,
                                                         ctypes.POINTER(c_f64), c_i32, c_vp, c_vp, c_vp, c_vp,
                                                         ctypes.c_uint64, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_i32, c_f64,
                                                         c_vp, c_vp, c_vp, c_vp]),
    "b200i_upload_simulate_rng_pipelined": (ctypes.c_int, [c_i64, c_i32, c_i64, ctypes.POINTER(SimConsts), c_vp, ctypes.c_uint32,
                                                           ctypes.POINTER(c_f64), c_i32, c_vp, c_vp, c_vp, c_vp,
                                                           ctypes.c_uint64, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_i32, c_f64,
                                                           c_vp, c_vp, c_vp, c_vp]),
    "b200i_gram_workspace_bytes": (c_i64, []),
    "b200i_theta_gram": (ctypes.c_int, [c_i64, c_i32, c_f64] + [c_vp] * 9),
    "b200i_stlsq_population": (ctypes.c_int, [c_vp, c_f64, c_f64, c_i32, c_vp, c_vp, c_vp]),
    "b200i_ode_rollout": (ctypes.c_int, [c_i64, c_i32, c_f64, c_i32, c_vp, c_vp, c_vp, c_vp, c_i32, c_f64, c_vp, c_vp]),
    "b200i_ode_rollout_f32": (ctypes.c_int, [c_i64, c_i32, c_f64, c_i32, c_vp, c_vp, c_vp, c_vp, c_i32, c_f64, c_vp, c_vp]),
    "b200i_treatment_codes": (ctypes.c_int, [c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "b200i_masked_se_workspace_bytes": (c_i64, []),
    "b200i_masked_se": (ctypes.c_int, [c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "b200i_sim_cf_one_step": (ctypes.c_int, [c_i64, c_i32, ctypes.POINTER(SimConsts)] + [c_vp] * 5 +
                              [c_i64, ctypes.POINTER(CfSource)] + [c_vp] * 6 +
                              [ctypes.POINTER(c_i64), ctypes.POINTER(c_i32), c_vp]),
    "b200i_sim_cf_treatment_seq": (ctypes.c_int, [c_i64, c_i32, c_i32, ctypes.POINTER(SimConsts)] + [c_vp] * 5 +
                                   [c_i64, ctypes.POINTER(CfSource)] + [c_vp] * 7 +
                                   [ctypes.POINTER(c_i64), ctypes.POINTER(c_i32), c_vp]),
    "b200i_stlsq_batched": (ctypes.c_int, [c_i64, c_i32, c_f64] + [c_vp] * 5 + [c_f64, c_f64, c_f64, c_i32, c_vp, c_vp]),
    "b200i_insite_bfgs": (ctypes.c_int, [c_i64, c_i32, c_f64, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_f64, c_f64,
                                         c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "b200i_insite_bfgs_joint": (ctypes.c_int, [c_i64, c_i32, c_f64, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_f64, c_f64,
                                               c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "b200i_cf_eval_one_step": (ctypes.c_int, [c_i64, c_i32, c_f64, c_i32] + [c_vp] * 6 + [c_i32, c_f64, c_vp, c_vp]),
    "b200i_cf_eval_treatment_seq": (ctypes.c_int, [c_i64, c_i32, c_i32, c_f64, c_i32] + [c_vp] * 7 + [c_i32, c_f64, c_vp, c_vp]),
    "b200i_insite_bfgs_prefix": (ctypes.c_int, [c_i64, c_i32, c_i32, c_f64, c_i32] + [c_vp] * 5 + [c_f64, c_f64, c_i32, c_i32] +
                                 [c_vp] * 4),
    "b200i_stlsq_prefix": (ctypes.c_int, [c_i64, c_i32, c_i32, c_f64] + [c_vp] * 5 + [c_f64, c_f64, c_f64, c_i32, c_vp, c_vp]),
    "b200i_ode_rollout_dts": (ctypes.c_int, [c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_i32, c_f64, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "b200i_stlsq_batched_dts": (ctypes.c_int, [c_i64, c_i32] + [c_vp] * 6 + [c_f64, c_f64, c_f64, c_i32, c_f64, c_vp, c_i32,
                                                                             c_i32, c_vp, c_vp]),
    "b200i_insite_bfgs_dts": (ctypes.c_int, [c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_f64, c_f64, c_i32,
                                             c_i32, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "b200i_poly_workspace_bytes": (c_i64, []),
    "b200i_poly_tsqr": (ctypes.c_int, [c_i64, c_i32, ctypes.c_double] + [c_vp] * 8),
    "b200i_poly_stlsq": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_double, c_i32, ctypes.c_double, c_vp, c_vp, c_vp]),
    "b200i_poly_rollout": (ctypes.c_int, [c_i64, c_i32, ctypes.c_double, c_i32] + [c_vp] * 4 + [ctypes.c_double, c_vp, c_vp]),
    "b200i_smooth_snippets": (ctypes.c_int, [c_i64, c_i32] + [c_vp] * 4 + [c_i32, c_vp, c_vp]),
    "b200i_theta_gram_dts": (ctypes.c_int, [c_i64, c_i32] + [c_vp] * 6 + [c_i32, c_vp, c_vp]),
    "b200i_expand_cf_one_step": (ctypes.c_int, [c_i64, c_i32] + [c_vp] * 5 + [c_i64, c_i64] + [c_vp] * 5 + [c_vp]),
    "b200i_expand_cf_treatment_seq": (ctypes.c_int, [c_i64, c_i32, c_i32] + [c_vp] * 6 + [c_i64, c_i64] + [c_vp] * 7 +
                                      [c_vp]),
}

_lib = None


def load():
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a). There is no CPU fallback for the INSITE hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().b200i_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with status {rc}: {msg}")
