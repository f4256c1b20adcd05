"""TEST INFRASTRUCTURE ONLY.  numpy / sklearn restatement of the reference SINDy-INSITE path.

The reference code (libs_m/ct/src/models/sindy.py, libs_m/ct/src/data/pkpd/utils.py) cannot be
imported here: it needs pysindy, jax, sympy2jax, hydra, pytorch-lightning, none of which are
installed (no network).  The arithmetic that lives in the un-vendored, *unpinned* dependency
``pysindy`` (setup/requirements.txt:1; API usage + log dates imply 1.7.x) is restated from its
published algorithm:

  PolynomialLibrary(degree=2, interaction_only=True)  -> features [1, x0, u0, x0*u0]
  FiniteDifference(order=1, is_uniform=True)          -> forward difference, last point backward
  STLSQ(threshold, alpha, max_iter=100)               -> loop as in the vendored copy
                                                         pkpd/utils.py:244-327 (ridge via
                                                         sklearn.linear_model.ridge_regression
                                                         :228) + pysindy's default unbias OLS refit
  call sites: sindy.py:185-213.

Parity pinning: tests/test_oracle.py checks this file against the reference's own committed run
log results/2_main_table/final_with_insite.txt:6 (16 population coefficients, 3 one-step RMSEs,
5 tau-step RMSEs for seed=1, gamma=2, 1000/100/100 patients) -- the only known-answer vectors the
reference holds for this path.

Other functions follow (file:line of the reference):
  process_data                   dataset.py:92-192
  process_sequential_test/multi  dataset.py:395-473, 533-552
  de_format_snippets             pkpd/utils.py:433-462, 543-554, 607-637
  equation_string                pkpd/utils.py:378-397
  rollout                        sindy.py:371-431 + pkpd/utils.py:68-90 (5 Euler sub-steps)
  masked_rmse / n_step_rmses     time_varying_model.py:236-313
  slice_autoregressive           sindy.py:729-733
  insite_objective / insite_bfgs sindy.py:587-631, 781-794 (scipy BFGS; iterate-level parity
                                 with jax's BFGS is UNPINNED)
"""
import numpy as np

STANDARD_DT = 10.0 / int(60)      # pkpd/utils.py:48-53
STEPS_FOR_DT = 5                  # pkpd/utils.py:40
TUMOUR_DEATH_THRESHOLD = 4 / 3 * np.pi * (13 / 2) ** 3


# ------------------------------------------------------------------------------------------------
# dataset transforms
# ------------------------------------------------------------------------------------------------
def process_data(sim, means, stds, treatment_mode='multiclass'):
    """dataset.py:92-192.  ``sim`` = simulator output dict; returns a NEW dict with the added keys."""
    d = dict(sim)
    cv = (sim['cancer_volume'] - means['cancer_volume']) / stds['cancer_volume']
    pt = (sim['patient_types'] - means['patient_types']) / stds['patient_types']
    pt = np.stack([pt for _ in range(cv.shape[1])], axis=1)
    chemo, radio, sl = sim['chemo_application'], sim['radio_application'], sim['sequence_lengths']
    treatments = np.concatenate([chemo[:, :-1, None], radio[:, :-1, None]], axis=-1)
    if treatment_mode == 'multiclass':
        code = (treatments[..., 0] == 1) * 1 + (treatments[..., 1] == 1) * 2
        known = ((treatments[..., 0] == 0) | (treatments[..., 0] == 1)) & \
                ((treatments[..., 1] == 0) | (treatments[..., 1] == 1))
        one_hot = np.zeros(treatments.shape[:2] + (4,))
        for c in range(4):
            one_hot[..., c] = ((code == c) & known) * 1.0
        d['prev_treatments'] = one_hot[:, :-1, :]
        d['current_treatments'] = one_hot
    else:
        d['prev_treatments'] = treatments[:, :-1, :]
        d['current_treatments'] = treatments
    cov = np.concatenate([cv[:, :-1, None], pt[:, :-1, None]], axis=-1)
    outputs = cv[:, 1:, None]
    active = np.zeros(outputs.shape)
    for i in range(sl.shape[0]):
        active[i, :int(sl[i]), :] = 1
    d['current_covariates'] = cov
    d['outputs'] = outputs
    d['active_entries'] = active
    d['unscaled_outputs'] = outputs * stds['cancer_volume'] + means['cancer_volume']
    d['prev_outputs'] = cov[:, :, :1]
    d['static_features'] = cov[:, 0, 1:]
    zero = np.zeros((cov.shape[0], 1, d['prev_treatments'].shape[-1]))
    d['prev_treatments'] = np.concatenate([zero, d['prev_treatments']], axis=1)
    scaling = {'input_means': np.array([means['cancer_volume'], means['patient_types'], 0.0, 0.0]),
               'inputs_stds': np.array([stds['cancer_volume'], stds['patient_types'], 1.0, 1.0]),
               'output_means': means['cancer_volume'], 'output_stds': stds['cancer_volume']}
    return d, scaling


def process_sequential_test(data, scaling, H):
    """dataset.py:395-473 (encoder_r=None): the last H steps of every row."""
    sl, outputs = data['sequence_lengths'], data['outputs']
    cur, prev = data['current_treatments'], data['prev_treatments'][:, 1:, :]
    cov = data['current_covariates']
    R = outputs.shape[0]
    o = np.zeros((R, H, outputs.shape[-1])); ct = np.zeros((R, H, cur.shape[-1]))
    pt = np.zeros((R, H, prev.shape[-1])); cc = np.zeros((R, H, cov.shape[-1]))
    for i in range(R):
        fl = int(sl[i]) - H
        pt[i] = prev[i, fl - 1:fl + H - 1, :]
        ct[i] = cur[i, fl:fl + H, :]
        o[i] = outputs[i, fl:fl + H, :]
        cc[i] = np.repeat([cov[i, fl - 1]], H, axis=0)
    return {'prev_treatments': pt, 'current_treatments': ct, 'current_covariates': cc,
            'prev_outputs': cc[:, :, :1], 'static_features': cc[:, 0, 1:], 'outputs': o,
            'sequence_lengths': np.full(R, float(H)), 'active_entries': np.ones((R, H, 1)),
            'unscaled_outputs': o * scaling['output_stds'] + scaling['output_means']}


# ------------------------------------------------------------------------------------------------
# population fit
# ------------------------------------------------------------------------------------------------
def de_format_snippets(data, scaling):
    """pkpd/utils.py:543-554 + :433-462 + :620-637 (joint=False, CANCER_SIM).

    Returns 4 lists (one per treatment) of (x (L,), u (L,)) constant-treatment snippets."""
    prev = data['prev_outputs'] * scaling['output_stds'] + scaling['output_means']
    static = data['static_features'] * scaling['inputs_stds'][1:2] + scaling['input_means'][1:2]
    cur = np.squeeze(data['current_treatments'])
    unscaled_outputs = np.squeeze(data['unscaled_outputs'])
    sl = data['sequence_lengths'].astype(np.int64)
    vol = np.concatenate((prev[:, 0].reshape(-1, 1), unscaled_outputs), axis=1)
    buckets = ([], [], [], [])
    for p in range(vol.shape[0]):
        tr, x, L, u = cur[p], vol[p], int(sl[p]), static[p, 0]
        cur_t, cur_x = [], []
        for i in range(L):
            if len(cur_t) >= 1 and (tr[i] != cur_t[-1]).any():
                cur_t.append(cur_t[-1]); cur_x.append(x[i])
                buckets[int(np.argmax(np.stack(cur_t).mean(0)))].append((np.array(cur_x), np.full(len(cur_x), u)))
                cur_t, cur_x = [tr[i]], [x[i]]
            else:
                cur_t.append(tr[i]); cur_x.append(x[i])
            if i == L - 1:
                cur_t.append(tr[i]); cur_x.append(x[i + 1])
                buckets[int(np.argmax(np.stack(cur_t).mean(0)))].append((np.array(cur_x), np.full(len(cur_x), u)))
    return buckets


def finite_difference_order1(x, dt):
    """pysindy FiniteDifference(order=1, is_uniform=True): forward, last point backward."""
    xd = np.empty_like(x)
    xd[:-1] = (x[1:] - x[:-1]) / dt
    xd[-1] = (x[-1] - x[-2]) / dt
    return xd


def savgol_w2_p1(x):
    """scipy.signal.savgol_filter(x, window_length=2, polyorder=1, axis=0) restated: pysindy's SmoothedFiniteDifference
    with the reference's smoother_kws (sindy.py:196-198).  savgol_coeffs of an even window evaluates its line fit at the
    half-sample position (pos = halflen - 0.5), i.e. weights [0.5, 0.5]; convolve1d places them on samples i and i+1;
    mode='interp' refits the first and the last sample through the outermost two samples, which returns them unchanged
    (up to lstsq rounding).  Pinned against scipy itself in tests/test_oracle.py."""
    x = np.asarray(x, dtype=np.float64)
    y = x.copy()
    y[1:-1] = 0.5 * x[1:-1] + 0.5 * x[2:]
    return y


def library_p4(x, u):
    """PolynomialLibrary(degree=2, interaction_only=True) on [x0, u0] -> [1, x0, u0, x0 u0]."""
    return np.stack([np.ones_like(x), x, u, x * u], axis=1)


POLY4_EXPONENTS = tuple((d - b, b) for d in range(5) for b in range(d + 1))     # (a, b) of x0^a u0^b, sklearn's order
POLY4_NAMES = ('1', 'x0', 'u0', 'x0^2', 'x0 u0', 'u0^2', 'x0^3', 'x0^2 u0', 'x0 u0^2', 'u0^3',
               'x0^4', 'x0^3 u0', 'x0^2 u0^2', 'x0 u0^3', 'u0^4')


def library_poly4(x, u):
    """PolynomialLibrary(degree=4, interaction_only=False) on [x0, u0] (sindy.py:185-186): 15 monomials, by total degree
    and lexicographic within a degree (sklearn PolynomialFeatures / pysindy get_feature_names order, POLY4_NAMES)."""
    return np.stack([x ** a * u ** b for a, b in POLY4_EXPONENTS], axis=1)


def design_matrices_poly4(snippets, dt=STANDARD_DT):
    th = np.concatenate([library_poly4(x, u) for x, u in snippets], axis=0)
    xd = np.concatenate([finite_difference_order1(x, dt) for x, _ in snippets], axis=0)
    return th, xd


def design_matrices(snippets, dt=STANDARD_DT, smoothed=False):
    """smoothed: SmoothedFiniteDifference -- pysindy (>= 1.7.4, FeatureLibrary.calc_trajectory) differentiates the
    smoothed trajectory and builds the library from it as well."""
    if smoothed:
        snippets = [(savgol_w2_p1(x), u) for x, u in snippets]
    th = np.concatenate([library_p4(x, u) for x, u in snippets], axis=0)
    xd = np.concatenate([finite_difference_order1(x, dt) for x, _ in snippets], axis=0)
    return th, xd


def stlsq_fit(theta, xdot, threshold, alpha, max_iter=100, unbias=True, scipy_lstsq=False):
    """pysindy STLSQ.  Loop: pkpd/utils.py:274-310 (vendored copy); ridge: :228; unbias: pysindy
    BaseOptimizer._unbias (LinearRegression without intercept on the final support).
    scipy_lstsq: un-bias with scipy.linalg.lstsq(X, y) directly -- what LinearRegression called in the scikit-learn of
    the reference's time (<= 1.2: cut-off = machine epsilon); the one installed here passes cond=tol=1e-6, which matters
    only for rank-deficient libraries (degree 4)."""
    import warnings
    from sklearn.linear_model import ridge_regression, LinearRegression
    n_feat = theta.shape[1]
    ind = np.ones(n_feat, dtype=bool)
    coef = np.linalg.lstsq(theta, xdot, rcond=None)[0]
    history = [coef.copy()]
    n_sel = n_feat
    for _ in range(max_iter):
        if np.count_nonzero(ind) == 0:
            coef = np.zeros(n_feat)
            break
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')            # scipy's ill-conditioning warning on the degree-4 library
            c_i = ridge_regression(theta[:, ind], xdot, alpha, tol=1e-6)
        c = np.zeros(n_feat)
        c[ind] = c_i
        big = np.abs(c) >= threshold
        c[~big] = 0
        coef, ind = c, big
        history.append(coef.copy())
        this, last = history[-1], history[-2] if len(history) > 1 else np.zeros_like(coef)
        if np.sum(ind) == n_sel or all(bool(a) == bool(b) for a, b in zip(this, last)):
            break
        n_sel = np.sum(ind)
    if unbias and np.any(ind):
        c = np.zeros(n_feat)
        if scipy_lstsq:
            import scipy.linalg
            c[ind] = scipy.linalg.lstsq(theta[:, ind], xdot)[0]
        else:
            c[ind] = LinearRegression(fit_intercept=False).fit(theta[:, ind], xdot).coef_
        coef = c
    return coef, ind


def fit_population_poly4(data, scaling, threshold=1e-3, alpha=0.5, dt=STANDARD_DT):
    """sindy.py:160-213 with ablation_more_complex_basis_functions -> joint_coefs (4,15), support (4,15) bool, and the
    explicit design matrices per treatment (for the statistics checks)."""
    buckets = de_format_snippets(data, scaling)
    coefs, sup, mats = np.zeros((4, 15)), np.zeros((4, 15), dtype=bool), []
    for a in range(4):
        th, xd = design_matrices_poly4(buckets[a], dt)
        coefs[a], sup[a] = stlsq_fit(th, xd, threshold, alpha, scipy_lstsq=True)
        mats.append((th, xd))
    return coefs, sup, mats


def rollout_unscaled_poly4(x0, codes, u, coefs, dt=STANDARD_DT, steps=STEPS_FOR_DT):
    """rollout_unscaled for dx/dt = sum_j coefs[code, j] x^a_j u^b_j (the sympy expression of the degree-4 model)."""
    R, W = codes.shape
    v = x0.astype(np.float64).copy()
    out = np.empty((R, W))
    h = dt / steps
    for k in range(W):
        c = coefs[codes[:, k]]
        for _ in range(steps):
            f = np.zeros(R)
            for j, (a, b) in enumerate(POLY4_EXPONENTS):
                f = f + c[:, j] * v ** a * u ** b
            v = v + f * h
        out[:, k] = v
    return out


def predictions_population_poly4(data, scaling, coefs):
    prev = np.squeeze(data['prev_outputs'] * scaling['output_stds'] + scaling['output_means'], -1)
    static = data['static_features'] * scaling['inputs_stds'][1:2] + scaling['input_means'][1:2]
    codes = np.argmax(data['current_treatments'], axis=-1)
    un = rollout_unscaled_poly4(prev[:, 0], codes, static[:, 0], effective_coefs(coefs))
    return ((un - scaling['output_means']) / scaling['output_stds'])[..., None]


def fit_population(data, scaling, threshold=1e-3, alpha=0.5, dt=STANDARD_DT, smoothed=False):
    """sindy.py:160-213, 332-336 -> joint_coefs (4,4), support (4,4) bool."""
    buckets = de_format_snippets(data, scaling)
    coefs, sup = np.zeros((4, 4)), np.zeros((4, 4), dtype=bool)
    stats = []
    for a in range(4):
        th, xd = design_matrices(buckets[a], dt, smoothed)
        coefs[a], sup[a] = stlsq_fit(th, xd, threshold, alpha)
        stats.append((len(buckets[a]), th.shape[0]))
    return coefs, sup, stats


# ---- joint model (one ODE over [x0, chemo, radio, static]; ablation "one_ode") ---------------------
JOINT_NAMES = ('1', 'x0', 'u0', 'u1', 'u2', 'x0*u0', 'x0*u1', 'x0*u2', 'u0*u1', 'u0*u2', 'u1*u2')


def de_format_joint(data, scaling):
    """pkpd/utils.py:543-554, :656-672 with process_sindy_training_data(joint=True) :493-497 for CANCER_SIM
    (sequence_lengths_offset = 0, sindy.py:160-171): one trajectory per patient,
    X = unscaled_outputs[:L] (the *outputs*, i.e. V[1:1+L]) and U = [current_treatments[:L] (multilabel:
    chemo, radio application), static repeated]."""
    static = data['static_features'] * scaling['inputs_stds'][1:2] + scaling['input_means'][1:2]
    cur = np.squeeze(data['current_treatments'])
    assert cur.shape[-1] == 2, "the joint model is configured with treatment_mode='multilabel'"
    out = np.squeeze(data['unscaled_outputs'])
    sl = data['sequence_lengths'].astype(np.int64)
    trajs = []
    for p in range(out.shape[0]):
        L = int(sl[p])
        trajs.append((out[p, :L], np.concatenate([cur[p, :L], np.full((L, 1), static[p, 0])], axis=1)))
    return trajs


def library_p11(x, U):
    """PolynomialLibrary(degree=2, interaction_only=True) on [x0, u0, u1, u2] (JOINT_NAMES order)."""
    u0, u1, u2 = U[:, 0], U[:, 1], U[:, 2]
    return np.stack([np.ones_like(x), x, u0, u1, u2, x * u0, x * u1, x * u2, u0 * u1, u0 * u2, u1 * u2], axis=1)


def design_matrices_joint(trajs, dt=STANDARD_DT, smoothed=False):
    if smoothed:
        trajs = [(savgol_w2_p1(x), U) for x, U in trajs]
    th = np.concatenate([library_p11(x, U) for x, U in trajs], axis=0)
    xd = np.concatenate([finite_difference_order1(x, dt) for x, _ in trajs], axis=0)
    return th, xd


def fit_population_joint(data, scaling, threshold=1e-3, alpha=0.5, dt=STANDARD_DT, smoothed=False):
    """sindy.py:185-204 with joint_model=True -> joint_coefs (1,11), support (11,), number of rows."""
    th, xd = design_matrices_joint(de_format_joint(data, scaling), dt, smoothed)
    coef, sup = stlsq_fit(th, xd, threshold, alpha)
    return coef[None, :], sup, th.shape[0]


def equation_string_joint(coefs):
    """pkpd/utils.py:386-391 + sindy.py:314."""
    s = ''
    for i, c in enumerate(coefs[0]):
        if np.abs(c) > 1e-3:
            s += f'+{c}*' + JOINT_NAMES[i]
    return f'Joint Model: x_dot = {s}'


def joint_to_per_treatment(coefs11):
    """The 11-term ODE restricted to a treatment (chemo, radio) in {0,1}^2 is a 4-term ODE in [1, x0, u2, x0*u2]:
    returns (4,4) indexed by the multiclass code chemo + 2*radio (what the rollout kernels take)."""
    c = np.asarray(coefs11, dtype=np.float64).reshape(-1)
    out = np.zeros((4, 4))
    for code in range(4):
        u0, u1 = float(code & 1), float(code >> 1)
        out[code] = [c[0] + c[2] * u0 + c[3] * u1 + c[8] * u0 * u1, c[1] + c[5] * u0 + c[6] * u1,
                     c[4] + c[9] * u0 + c[10] * u1, c[7]]
    return out


def predictions_population_joint(data, scaling, coefs11, dt=STANDARD_DT, steps=STEPS_FOR_DT):
    """_get_non_fine_tuned_predictions (sindy.py:371-431) with the joint closure (:316-318): Euler rollout of
    the 11-term expression, treatments = the two multilabel columns."""
    prev = np.squeeze(data['prev_outputs'] * scaling['output_stds'] + scaling['output_means'], -1)
    static = data['static_features'] * scaling['inputs_stds'][1:2] + scaling['input_means'][1:2]
    cur = np.squeeze(data['current_treatments']).astype(np.int64).astype(np.float64)   # :396 astype(int64)
    c = effective_coefs(np.asarray(coefs11, dtype=np.float64).reshape(-1))
    R, W, _ = cur.shape
    v = prev[:, 0].astype(np.float64).copy()
    u2 = static[:, 0]
    out = np.empty((R, W))
    h = dt / steps
    for k in range(W):
        u0, u1 = cur[:, k, 0], cur[:, k, 1]
        for _ in range(steps):
            f = (c[0] + c[1] * v + c[2] * u0 + c[3] * u1 + c[4] * u2 + c[5] * v * u0 + c[6] * v * u1 + c[7] * v * u2
                 + c[8] * u0 * u1 + c[9] * u0 * u2 + c[10] * u1 * u2)
            v = v + f * h
        out[:, k] = v
    return ((out - scaling['output_means']) / scaling['output_stds'])[..., None]


def equation_string(coefs, names=('1', 'x0', 'u0', 'x0*u0')):
    """pkpd/utils.py:386-391 + sindy.py:295."""
    parts = []
    for a in range(coefs.shape[0]):
        s = ''
        for i, c in enumerate(coefs[a]):
            if np.abs(c) > 1e-3:
                s += f'+{c}*' + names[i]
        parts.append(f'Treatment {a}: x_dot = {s}')
    return ' | '.join(parts)


# ------------------------------------------------------------------------------------------------
# rollout + metrics
# ------------------------------------------------------------------------------------------------
def odeint_euler(func, y0, t, *args, hmax=np.inf, steps=STEPS_FOR_DT):
    """odeint of the reference (pkpd/utils.py:68-90) for an arbitrary time grid t: dts = diff(t); when hmax < dts[0]
    every interval is split into `steps` Euler sub-steps of dts[k] / steps (odeint_high_resolution_euler), else one
    Euler step per interval; y <- y + func(y, h, *args) * h (scan_func).  Returns y at every grid point."""
    t = np.asarray(t, dtype=np.float64)
    dts = np.diff(t)
    y = np.asarray(y0, dtype=np.float64)
    out = [y]
    fine = hmax < dts[0]
    for d in dts:
        if fine:
            h = d / steps
            for _ in range(steps):
                y = y + func(y, h, *args) * h
        else:
            y = y + func(y, d, *args) * d
        out.append(y)
    return np.stack(out, axis=0)


def rollout_unscaled(x0, codes, u, coefs, dt=STANDARD_DT, steps=STEPS_FOR_DT, dts=None):
    """sindy.py:413-431 / :767-778 with pkpd/utils.py:68-90: open loop, 5 Euler sub-steps per
    interval.  x0 (R,), codes (R,W) int treatment index (argmax of the one-hot), u (R,),
    coefs (4,4) or per-row (R,4,4).  Returns (R,W).
    dts: interval lengths (W,) or (R,W) of an irregular grid: interval k is integrated as odeint over [0, dts[k]]."""
    R, W = codes.shape
    v = x0.astype(np.float64).copy()
    out = np.empty((R, W))
    h = dt / steps
    rows = np.arange(R)
    for k in range(W):
        if dts is not None:
            h = (dts[k] if np.ndim(dts) == 1 else dts[:, k]) / steps
        c = coefs[codes[:, k]] if coefs.ndim == 2 else coefs[rows, codes[:, k]]
        for _ in range(steps):
            v = v + (c[:, 0] * 1 + c[:, 1] * v + c[:, 2] * u + c[:, 3] * (v * u)) * h
        out[:, k] = v
    return out


def effective_coefs(coefs):
    """The sympy expression only keeps terms with |c| > 1e-3 (pkpd/utils.py:387-391)."""
    return np.where(np.abs(coefs) > 1e-3, coefs, 0.0)


def predictions_population(data, scaling, coefs):
    """_get_non_fine_tuned_predictions sindy.py:371-431 -> scaled predictions (R,W,1)."""
    prev = np.squeeze(data['prev_outputs'] * scaling['output_stds'] + scaling['output_means'], -1)
    static = data['static_features'] * scaling['inputs_stds'][1:2] + scaling['input_means'][1:2]
    codes = np.argmax(data['current_treatments'], axis=-1)
    un = rollout_unscaled(prev[:, 0], codes, static[:, 0], effective_coefs(coefs))
    return ((un - scaling['output_means']) / scaling['output_stds'])[..., None]


def masked_rmse(pred_scaled, data, scaling, norm_const=TUMOUR_DEATH_THRESHOLD, one_step_counterfactual=True):
    """time_varying_model.py:236-283 (unscale=True, percentage=True) -> (orig, all, last)."""
    un = pred_scaled * scaling['output_stds'] + scaling['output_means']
    act = data['active_entries']
    mse = ((un - data['unscaled_outputs']) ** 2) * act
    orig = (mse.sum(0).sum(-1) / act.sum(0).sum(-1)).mean()
    rmse_orig = np.sqrt(orig) / norm_const * 100.0
    rmse_all = np.sqrt(mse.sum() / act.sum()) / norm_const * 100.0
    if not one_step_counterfactual:
        return rmse_orig, rmse_all
    R, _, od = act.shape
    last = act - np.concatenate([act[:, 1:, :], np.zeros((R, 1, od))], axis=1)
    mse_last = (((un - data['unscaled_outputs']) ** 2) * last).sum() / last.sum()
    return rmse_orig, rmse_all, np.sqrt(mse_last) / norm_const * 100.0


def slice_autoregressive(pred_scaled, sequence_lengths, H):
    """sindy.py:729-733: dynamic_slice(preds, (i, max(1, sl-H), 0), (1,H,1)) (start is clamped
    so that the slice fits, as lax.dynamic_slice does)."""
    R, W, _ = pred_scaled.shape
    sl = sequence_lengths.astype(np.int64)
    lo = np.clip(np.maximum(1, sl - H), 0, W - H)
    idx = lo[:, None] + np.arange(H)[None, :]
    return pred_scaled[np.arange(R)[:, None], idx, :]


def n_step_rmses(pred_seq_scaled, data_seq, scaling, norm_const=TUMOUR_DEATH_THRESHOLD):
    """time_varying_model.py:285-313 -> (H,) percentages."""
    un = pred_seq_scaled * scaling['output_stds'] + scaling['output_means']
    mse = ((un - data_seq['unscaled_outputs']) ** 2) * data_seq['active_entries']
    nan_idx = np.unique(np.where(np.isnan(data_seq['outputs']))[0])
    keep = np.setdiff1d(np.arange(un.shape[0]), nan_idx)
    mse_orig = mse[keep].sum(0).sum(-1) / data_seq['active_entries'][keep].sum(0).sum(-1)
    return np.sqrt(mse_orig) / norm_const * 100.0


# ------------------------------------------------------------------------------------------------
# individualisation (INSITE)
# ------------------------------------------------------------------------------------------------
def insite_objective(theta_flat, x, codes, u, n_fit, theta0_flat, lam, norm, dt=STANDARD_DT, steps=STEPS_FOR_DT,
                     with_grad=True, dts=None):
    """f_to_min_func (sindy.py:781-794) for one row + its gradient by forward sensitivities.
    x: un-scaled prev_outputs (W,), codes (W,) treatment index, n_fit = sequence_length - projection_horizon."""
    theta = np.asarray(theta_flat, dtype=np.float64).reshape(4, 4)
    theta0 = np.asarray(theta0_flat, dtype=np.float64).reshape(4, 4)
    mask = (np.abs(theta0) > 1e-3).astype(np.float64)          # coef_sparse_mask, sindy.py:589
    c = theta * mask
    h = dt / steps
    v = float(x[0])
    s = np.zeros((4, 4))
    acc = 0.0
    gacc = np.zeros((4, 4))
    for k in range(int(n_fit)):
        a = int(codes[k])
        c0, c1, c2, c3 = c[a]
        if dts is not None:
            h = dts[k] / steps
        for _ in range(steps):
            vu = v * u
            f = ((c0 + c1 * v) + c2 * u) + c3 * vu
            if with_grad:
                dfdv = c1 + c3 * u
                df = np.zeros((4, 4))
                df[a] = np.array([1.0, v, u, vu]) * mask[a]
                s = s + h * (dfdv * s + df)
            v = v + h * f
        r = x[k + 1] - v
        acc += r * r
        if with_grad:
            gacc += -2.0 * r * s
    mse = acc / n_fit
    diff = theta - theta0
    val = mse / norm + lam * np.mean(diff ** 2)
    if not with_grad:
        return val
    grad = gacc / n_fit / norm + lam * 2.0 * diff / 16.0
    return val, grad.reshape(-1)


def insite_objective_joint(theta11, x, codes, u, n_fit, theta0_11, lam, norm, dt=STANDARD_DT, steps=STEPS_FOR_DT,
                            with_grad=True):
    """f_to_min_func (sindy.py:781-794) for the joint model (pred_dy_dt of :503-517: one 11-term expression in
    x0 = volume, u0 = chemo, u1 = radio application, u2 = static feature) + gradient by forward sensitivities.
    codes = chemo + 2*radio per step."""
    theta = np.asarray(theta11, dtype=np.float64).reshape(11)
    theta0 = np.asarray(theta0_11, dtype=np.float64).reshape(11)
    mask = (np.abs(theta0) > 1e-3).astype(np.float64)
    c = theta * mask
    h = dt / steps
    v = float(x[0])
    s = np.zeros(11)
    acc = 0.0
    gacc = np.zeros(11)
    for k in range(int(n_fit)):
        u0, u1 = float(int(codes[k]) & 1), float(int(codes[k]) >> 1)
        for _ in range(steps):
            basis = np.array([1.0, v, u0, u1, u, v * u0, v * u1, v * u, u0 * u1, u0 * u, u1 * u])
            f = float(np.dot(c, basis))
            if with_grad:
                dfdv = c[1] + c[5] * u0 + c[6] * u1 + c[7] * u
                s = s + h * (dfdv * s + basis * mask)
            v = v + h * f
        r = x[k + 1] - v
        acc += r * r
        if with_grad:
            gacc += -2.0 * r * s
    diff = theta - theta0
    val = acc / n_fit / norm + lam * np.mean(diff ** 2)
    if not with_grad:
        return val
    return val, gacc / n_fit / norm + lam * 2.0 * diff / 11.0


def insite_bfgs_row_joint(x, codes, u, seq_len, ph, theta0_11, lam, gtol=1e-12, maxiter=3200):
    """_fine_tuning_inner for the joint model with scipy's BFGS (UNPINNED at the iterate level, like insite_bfgs_row)."""
    from scipy.optimize import minimize
    n_fit = int(min(seq_len - ph, len(x) - 1))
    theta0 = np.asarray(theta0_11, dtype=np.float64).reshape(11)
    if n_fit <= 0:
        return theta0.copy(), 0.0, 0.0
    start = insite_objective_joint(theta0, x, codes, u, n_fit, theta0, lam, 1.0, with_grad=False)
    norm = 2.5 * start
    if not (norm > 0) or not np.isfinite(norm):
        return theta0.copy(), 0.0, 0.0
    fun = lambda t: insite_objective_joint(t, x, codes, u, n_fit, theta0, lam, norm)
    res = minimize(fun, theta0.copy(), jac=True, method='BFGS', options={'gtol': gtol, 'maxiter': maxiter})
    f0 = fun(theta0)[0]
    return (res.x if res.fun <= f0 else theta0.copy()), f0, min(res.fun, f0)


def insite_bfgs_row(x, codes, u, seq_len, ph, theta0, lam, gtol=1e-12, maxiter=3200):
    """_fine_tuning_inner (sindy.py:587-631) with scipy's BFGS standing in for jax's (UNPINNED: the two
    differ at the iterate level).  Returns (theta (4,4), f0, f_end)."""
    from scipy.optimize import minimize
    W = len(x)
    n_fit = min(int(seq_len) - ph, W - 1)
    t0 = np.asarray(theta0, dtype=np.float64).reshape(-1)
    if n_fit <= 0:
        return t0.reshape(4, 4).copy(), 0.0, 0.0
    start = insite_objective(t0, x, codes, u, n_fit, t0, lam, 1.0, with_grad=False)
    norm = 2.5 * start
    fun = lambda th: insite_objective(th, x, codes, u, n_fit, t0, lam, norm)
    f0 = fun(t0)[0]
    res = minimize(fun, t0, jac=True, method='BFGS', options={'gtol': gtol, 'maxiter': maxiter})
    th = res.x if res.fun <= f0 else t0
    return th.reshape(4, 4), f0, min(res.fun, f0)


def ridge_prior_row(x, codes, u, n_fit, prior, lam, threshold=1e-3, support_tol=1e-3, max_iter=10, dt=STANDARD_DT,
                    dts=None):
    """Batched per-row estimator of the north star (K5b): per treatment, mean-normalised normal equations of
    the row's snippets, ridge shrunk to the population coefficients on the population support, thresholded."""
    prior = np.asarray(prior, dtype=np.float64).reshape(4, 4)
    G = np.zeros((4, 4, 4)); b = np.zeros((4, 4)); cnt = np.zeros(4)
    W = len(x)
    for k in range(int(n_fit)):
        a = int(codes[k]); a1 = int(codes[min(k + 1, W - 1)])
        xd = (x[k + 1] - x[k]) / (dt if dts is None else dts[k])
        pts = [x[k]] + ([x[k + 1]] if (k == n_fit - 1 or a1 != a) else [])
        for xv in pts:
            th = np.array([1.0, xv, u, xv * u])
            G[a] += np.outer(th, th); b[a] += th * xd; cnt[a] += 1
    out = prior.copy()
    for a in range(4):
        ind = np.abs(prior[a]) > support_tol
        if cnt[a] == 0 or not ind.any():
            continue
        Ga, ba = G[a] / cnt[a], b[a] / cnt[a]
        c = np.zeros(4)
        for _ in range(max_iter):
            c = np.zeros(4)
            c[ind] = np.linalg.solve(Ga[np.ix_(ind, ind)] + lam * np.eye(ind.sum()), ba[ind] + lam * prior[a][ind])
            big = ind & (np.abs(c) >= threshold)
            if np.array_equal(big, ind):
                break
            ind = big
            if not ind.any():
                c = np.zeros(4)
                break
        out[a] = c
    return out


def normal_equations_irregular(vol, chemo, radio, seq_len, static, dts):
    """SURVEY.md App. B on an irregular grid: per treatment G = Theta^T Theta, b = Theta^T xdot, sample counts, with
    xdot_k = (x[k+1] - x[k]) / dts[k] (pysindy FiniteDifference(order=1) on a time array: forward differences, the last
    point of a snippet backward).  vol (N,T), dts (T-1,) or (N,T-1).  Explicit loops: small cohorts only."""
    N, T = vol.shape
    G = np.zeros((4, 4, 4)); b = np.zeros((4, 4)); cnt = np.zeros(4)
    for i in range(N):
        L = min(int(seq_len[i]), T - 1)
        u = static[i]
        a = (chemo[i] != 0).astype(int) + 2 * (radio[i] != 0).astype(int)
        d = dts if np.ndim(dts) == 1 else dts[i]
        for k in range(L):
            xd = (vol[i, k + 1] - vol[i, k]) / d[k]
            pts = [vol[i, k]] + ([vol[i, k + 1]] if (k == L - 1 or a[k + 1] != a[k]) else [])
            for xv in pts:
                th = np.array([1.0, xv, u, xv * u])
                G[a[k]] += np.outer(th, th); b[a[k]] += th * xd; cnt[a[k]] += 1
    return G, b, cnt


def lsq_initial_mask(theta, xdot, initial_guess, threshold, alpha, max_iter=100, unbias=False):
    """The reference's dormant per-patient optimiser LSQIntialMask._reduce (pkpd/utils.py:244-327, vendored copy of
    pysindy's STLSQ with a warm start): the initial guess only sets the initial support ind = |guess| > 1e-14
    (:251-253); every iteration ridge-regresses on the support (sklearn ridge_regression, :228) and thresholds (:213-219);
    stop when no feature was dropped w.r.t. the initial support or the pattern repeats (:308).  unbias=True adds
    pysindy's OLS refit on the final support (BaseOptimizer._unbias; the reference switches it off when the refit
    overflows, pkpd_simulation.py:795-797)."""
    from sklearn.linear_model import ridge_regression, LinearRegression
    n_feat = theta.shape[1]
    ind = np.abs(np.asarray(initial_guess, dtype=np.float64)) > 1e-14
    n_sel0 = int(ind.sum())
    coef = np.zeros(n_feat)
    history = [np.asarray(initial_guess, dtype=np.float64).copy()]
    for _ in range(max_iter):
        if not ind.any():
            coef = np.zeros(n_feat)
            break
        c = np.zeros(n_feat)
        c[ind] = ridge_regression(theta[:, ind], xdot, alpha, tol=1e-6)
        big = np.abs(c) >= threshold
        c[~big] = 0
        coef, ind = c, big
        history.append(coef.copy())
        if ind.sum() == n_sel0 or all(bool(a) == bool(b) for a, b in zip(history[-1], history[-2])):
            break
    if unbias and ind.any():
        c = np.zeros(n_feat)
        c[ind] = LinearRegression(fit_intercept=False).fit(theta[:, ind], xdot).coef_
        coef = c
    return coef, ind


def row_design_matrices(x, codes, u, n_fit, dt=STANDARD_DT):
    """Per-treatment design matrices of one row's fit window under the live path's rules (constant-treatment snippets,
    order-1 finite differences, last point of a snippet backward; App. B): list of (Theta (m,4), xdot (m,)) for a = 0..3."""
    W = len(x)
    th = [[] for _ in range(4)]; xd = [[] for _ in range(4)]
    for k in range(int(n_fit)):
        a = int(codes[k]); a1 = int(codes[min(k + 1, W - 1)])
        d = (x[k + 1] - x[k]) / dt
        for xv in [x[k]] + ([x[k + 1]] if (k == n_fit - 1 or a1 != a) else []):
            th[a].append([1.0, xv, u, xv * u]); xd[a].append(d)
    return [(np.array(th[a]).reshape(-1, 4), np.array(xd[a])) for a in range(4)]
