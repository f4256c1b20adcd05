// stlsq.cuh -- sequentially thresholded least squares on small normal equations (P <= 4 features),
// everything in registers.  Semantics: pysindy STLSQ as vendored at pkpd/utils.py:244-327 (ridge by
// Cholesky on G + alpha*I, as sklearn.linear_model.ridge_regression does for dense tall inputs,
// :228; hard threshold |c| >= threshold, :213-219; stop when no feature was dropped relative to the
// initial support or the support repeats, :308) followed by pysindy's default unbias step
// (ordinary least squares on the final support).
#pragma once
#include "common.cuh"

namespace b200i {

// Solve (G[S,S] + ridge*I) c[S] = b[S] (+ ridge*prior[S]) for the selected features S = mask by
// Cholesky with Jacobi (diagonal) scaling; unselected coefficients are 0.  The system is kept 4x4:
// unselected rows/columns are replaced by the identity, so every index is static (registers only).
// Returns false if the matrix is not positive definite.
__host__ __device__ inline bool solve_spd4(const double (&G)[4][4], const double (&b)[4], unsigned mask, double ridge,
                                           const double *prior, double (&c)[4])
{
    double A[4][4], r[4], sc[4];
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool si = (mask >> i) & 1u;
        const double d = si ? G[i][i] + ridge : 1.0;
        ok = ok && (d > 0.0);
        sc[i] = si ? 1.0 / sqrt(d > 0.0 ? d : 1.0) : 1.0;
        r[i] = si ? (b[i] + (prior ? ridge * prior[i] : 0.0)) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool sel = ((mask >> i) & 1u) && ((mask >> j) & 1u);
            const double g = G[i][j] + ((i == j) ? ridge : 0.0);
            A[i][j] = sel ? g * sc[i] * sc[j] : ((i == j) ? 1.0 : 0.0);
        }
        r[i] *= sc[i];
    }
    // Cholesky A = L L^T, lower triangle in place
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double d = A[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= A[j][k] * A[j][k];
        ok = ok && (d > 0.0);
        d = sqrt(d > 0.0 ? d : 1.0);
        A[j][j] = d;
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            double v = A[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= A[i][k] * A[j][k];
            A[i][j] = v / d;
        }
    }
    double y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double v = r[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v -= A[i][k] * y[k];
        y[i] = v / A[i][i];
    }
#pragma unroll
    for (int i = 3; i >= 0; --i) {
        double v = y[i];
#pragma unroll
        for (int k = i + 1; k < 4; ++k) v -= A[k][i] * y[k];
        y[i] = v / A[i][i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = ((mask >> i) & 1u) ? y[i] * sc[i] : 0.0;
    return ok;
}

// full STLSQ + unbias on one 4-feature problem.  Returns the final support mask.
__host__ __device__ inline unsigned stlsq4(const double (&G)[4][4], const double (&b)[4], double threshold, double alpha,
                                           int max_iter, unsigned init_mask, double (&coef)[4])
{
    unsigned ind = init_mask;
    int n_selected0 = 0;
    for (int j = 0; j < 4; ++j) n_selected0 += (ind >> j) & 1u;
    unsigned prev_pattern = 0xFu;  // history_[0] is the dense OLS initial guess (all non-zero)
    for (int j = 0; j < 4; ++j) coef[j] = 0.0;
    for (int it = 0; it < max_iter; ++it) {
        if (ind == 0) {
            for (int j = 0; j < 4; ++j) coef[j] = 0.0;
            break;
        }
        double c[4];
        if (!solve_spd4(G, b, ind, alpha, nullptr, c)) break;
        unsigned big = 0;
        for (int j = 0; j < 4; ++j) {
            if (((ind >> j) & 1u) && fabs(c[j]) >= threshold) big |= 1u << j;
            else c[j] = 0.0;
        }
        for (int j = 0; j < 4; ++j) coef[j] = c[j];
        ind = big;
        unsigned pattern = 0;
        for (int j = 0; j < 4; ++j) pattern |= (coef[j] != 0.0 ? 1u : 0u) << j;
        int n_sel = 0;
        for (int j = 0; j < 4; ++j) n_sel += (ind >> j) & 1u;
        const bool no_change = (pattern == prev_pattern);
        prev_pattern = pattern;
        if (n_sel == n_selected0 || no_change) break;
    }
    if (ind != 0) {  // unbias: OLS on the support
        double c[4];
        if (solve_spd4(G, b, ind, 0.0, nullptr, c))
            for (int j = 0; j < 4; ++j) coef[j] = c[j];
    }
    return ind;
}

// unpack the 15 packed statistics of one treatment into G (symmetric) and b
__host__ __device__ inline void unpack_gram(const double *g, double (&G)[4][4], double (&b)[4])
{
    int k = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j) {
            G[i][j] = g[k];
            G[j][i] = g[k];
            ++k;
        }
    for (int i = 0; i < 4; ++i) b[i] = g[10 + i];
}

}  // namespace b200i
