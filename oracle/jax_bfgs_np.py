"""TEST INFRASTRUCTURE ONLY.  numpy restatement of ``jax.scipy.optimize.minimize(fun, x0, method='BFGS')``.

The reference individualises every test row with it (libs_m/ct/src/models/sindy.py:627,
``minimize(f_to_min, reduced_coefs.reshape(-1), method='BFGS', tol=1e-12)``) and replaces the coefficients of rows
whose result has ``status == 3`` by the population's (:628-631).  jax is an un-vendored, unpinned dependency
(setup/install.sh:4) that is not installed here, so its published algorithm is restated
(jax/_src/scipy/optimize/{minimize,bfgs,line_search}.py, jax 0.4.x) and anchored on the reference's own committed
logs: with these semantics the product path reproduces the INSITE line of results/2_main_table/final_with_insite.txt
(:2362, no fallback: the revision of 2023-05-14) to 5e-15 and of results/ablation/one_ode/...txt (:6, with the fallback:
2023-05-16) to 2e-6 relative (tests/test_gpu_insite.py); this file is checked against the CUDA kernel row by row.

The semantics that matter
  * ``minimize`` does not forward ``tol``: minimize_bfgs runs with gtol = 1e-5 (infinity norm), maxiter = 200 * n;
  * line_search: first trial min(1, 1.01 * 2 (f_k - f_{k-1}) / dphi_0) (f_{-1} = f_0 + |g_0| / 2), then doubling, at most
    ``line_search_maxiter`` = 10 trials; a trial that violates sufficient decrease (or does not decrease w.r.t. the
    previous trial) starts zoom(a_{i-1}, a_i), one with a non-negative slope starts zoom(a_i, a_{i-1});
  * _zoom: trial point by cubic (after the first iteration), quadratic or bisection with the bounds checks on the
    SIGNED width dalpha = a_hi - a_lo; it FAILS when dalpha <= 1e-10 (float64) -- at once for a reversed bracket -- or
    after 30 iterations; the iteration that notices the failure still evaluates its trial point;
  * a failed line search ends BFGS (status 2 + line-search status: 3 = zoom failed, 5 = bracketing exhausted) and leaves
    x + a p, a = the zoom's a_star (its initial value 1 unless a trial satisfied both Wolfe conditions).
"""
import numpy as np


def _cubicmin(a, fa, fpa, b, fb, c, fc):
    with np.errstate(all='ignore'):
        C = fpa
        db, dc = b - a, c - a
        denom = (db * dc) ** 2 * (db - dc)
        d1 = np.array([[dc ** 2, -db ** 2], [-dc ** 3, db ** 3]])
        d2 = np.array([fb - fa - C * db, fc - fa - C * dc])
        A, B = d1.dot(d2) / denom
        radical = B * B - 3.0 * A * C
        return a + (-B + np.sqrt(radical)) / (3.0 * A)


def _quadmin(a, fa, fpa, b, fb):
    with np.errstate(all='ignore'):
        db = b - a
        B = (fb - fa - fpa * db) / (db ** 2)
        return a - fpa / (2.0 * B)


def _zoom(phi, wolfe_one, wolfe_two, a_lo, phi_lo, dphi_lo, a_hi, phi_hi, dphi_hi, g_0):
    """Returns (failed, a_star, phi_star, g_star, nfev)."""
    a_rec, phi_rec = (a_lo + a_hi) / 2.0, (phi_lo + phi_hi) / 2.0
    a_star, phi_star, g_star = 1.0, phi_lo, g_0
    done = failed = False
    j = nfev = 0
    while not done and not failed:
        dalpha = a_hi - a_lo
        a, b = min(a_hi, a_lo), max(a_hi, a_lo)
        cchk, qchk = 0.2 * dalpha, 0.1 * dalpha
        failed = failed or (dalpha <= 1e-10)
        a_j_cubic = _cubicmin(a_lo, phi_lo, dphi_lo, a_hi, phi_hi, a_rec, phi_rec)
        use_cubic = (j > 0) and (a_j_cubic > a + cchk) and (a_j_cubic < b - cchk)
        a_j_quad = _quadmin(a_lo, phi_lo, dphi_lo, a_hi, phi_hi)
        use_quad = (not use_cubic) and (a_j_quad > a + qchk) and (a_j_quad < b - qchk)
        a_j = a_j_cubic if use_cubic else (a_j_quad if use_quad else (a_lo + a_hi) / 2.0)
        phi_j, dphi_j, g_j = phi(a_j)
        nfev += 1
        hi_to_j = wolfe_one(a_j, phi_j) or (phi_j >= phi_lo)
        star_to_j = wolfe_two(dphi_j) and not hi_to_j
        hi_to_lo = (dphi_j * (a_hi - a_lo) >= 0.0) and not hi_to_j and not star_to_j
        lo_to_j = not hi_to_j and not star_to_j
        old = (a_lo, phi_lo, dphi_lo, a_hi, phi_hi, dphi_hi)
        if hi_to_j:
            a_rec, phi_rec = old[3], old[4]
            a_hi, phi_hi, dphi_hi = a_j, phi_j, dphi_j
        if star_to_j:
            done = True
            a_star, phi_star, g_star = a_j, phi_j, g_j
        if hi_to_lo:
            a_hi, phi_hi, dphi_hi = old[0], old[1], old[2]
            a_rec, phi_rec = old[3], old[4]
        if lo_to_j:
            a_rec, phi_rec = old[0], old[1]
            a_lo, phi_lo, dphi_lo = a_j, phi_j, dphi_j
        j += 1
        failed = failed or (j >= 30)
    return failed, a_star, phi_star, g_star, nfev


def line_search(fg, xk, pk, old_fval, old_old_fval, gfk, c1=1e-4, c2=0.9, maxiter=10):
    """Returns dict(failed, status, a_k, f_k, g_k)."""
    def phi(t):
        f, g = fg(xk + t * pk)
        return f, float(np.dot(g, pk)), g
    phi_0, dphi_0 = old_fval, float(np.dot(gfk, pk))
    with np.errstate(all='ignore'):
        cand = 1.01 * 2 * (phi_0 - old_old_fval) / dphi_0
    start = 1.0 if cand > 1 else cand
    wolfe_one = lambda a_i, phi_i: phi_i > phi_0 + c1 * a_i * dphi_0
    wolfe_two = lambda dphi_i: abs(dphi_i) <= -c2 * dphi_0
    done = failed = False
    i, a_i1, phi_i1, dphi_i1 = 1, 0.0, phi_0, dphi_0
    a_star, phi_star, g_star = 0.0, phi_0, gfk
    while not done and i <= maxiter and not failed:
        a_i = start if i == 1 else a_i1 * 2.0
        phi_i, dphi_i, g_i = phi(a_i)
        to_zoom1 = wolfe_one(a_i, phi_i) or (phi_i >= phi_i1 and i > 1)
        to_i = wolfe_two(dphi_i) and not to_zoom1
        to_zoom2 = (dphi_i >= 0.0) and not to_zoom1 and not to_i
        if to_zoom1:
            zf, a_star, phi_star, g_star, _ = _zoom(phi, wolfe_one, wolfe_two, a_i1, phi_i1, dphi_i1, a_i, phi_i, dphi_i, gfk)
            done, failed = True, failed or zf
        if to_i:
            done = True
            a_star, phi_star, g_star = a_i, phi_i, g_i
        if to_zoom2:
            zf, a_star, phi_star, g_star, _ = _zoom(phi, wolfe_one, wolfe_two, a_i, phi_i, dphi_i, a_i1, phi_i1, dphi_i1, gfk)
            done, failed = True, failed or zf
        i += 1
        a_i1, phi_i1, dphi_i1 = a_i, phi_i, dphi_i
    status = 1 if failed else (3 if i > maxiter else 0)
    return dict(failed=failed or not done, status=status, a_k=a_star, f_k=phi_star, g_k=g_star)


def minimize_bfgs(fg, x0, gtol=1e-5, maxiter=None, line_search_maxiter=10):
    """fg(x) -> (f, grad).  Returns dict(x, fun, status, nit): status 0 converged, 1 maxiter, 3 zoom failed,
    5 bracketing exhausted (2 + line-search status)."""
    x = np.asarray(x0, dtype=np.float64).copy()
    d = x.shape[0]
    maxiter = d * 200 if maxiter is None else maxiter
    f, g = fg(x)
    H = np.eye(d)
    old_old = f + np.linalg.norm(g) / 2
    converged = np.max(np.abs(g)) < gtol
    failed, k, ls_status = False, 0, 0
    while not converged and not failed and k < maxiter:
        p = -H.dot(g)
        ls = line_search(fg, x, p, f, old_old, g, maxiter=line_search_maxiter)
        failed, ls_status = ls['failed'], ls['status']
        s = ls['a_k'] * p
        x_new, f_new, g_new = x + s, ls['f_k'], ls['g_k']
        y = g_new - g
        with np.errstate(all='ignore'):
            rho = 1.0 / np.dot(y, s)
            w = np.eye(d) - rho * np.outer(s, y)
            H_new = w.dot(H).dot(w.T) + rho * np.outer(s, s)
        if np.isfinite(rho):
            H = H_new
        converged = np.max(np.abs(g_new)) < gtol
        old_old = f
        x, f, g, k = x_new, f_new, g_new, k + 1
    status = 0 if converged else (1 if k == maxiter else (2 + ls_status if failed else -1))
    return dict(x=x, fun=f, status=status, nit=k)
