// insite_fit.cu -- individualisation of the population ODE (the "I" of INSITE).
//
// Two estimators behind the same row interface (x = un-scaled prev_outputs (R,W), treatment codes
// (R,W), fit window = the first n_fit = sequence_length - projection_horizon transitions of the row,
// as create_mask(...) in f_to_min_func, sindy.py:786):
//
//  K5b  b200i_stlsq_batched   -- batched per-row sequentially-thresholded ridge regression shrunk to
//       the population coefficients on the population support (the BASELINE north-star estimator;
//       lineage: the reference's dormant per-patient STLSQ with warm start, pkpd_simulation.py:791-836
//       + LSQIntialMask, pkpd/utils.py:96-335).  One thread per row, 4x4 Cholesky in registers.
//
//  K7   b200i_insite_bfgs     -- the reference's live path (sindy.py:587-631, 781-794): BFGS over the
//       16 coefficients of  mean_{k<n_fit}(x[k+1]-xhat[k+1](theta*mask))^2 / (2.5*mse(theta0))
//       + lam*mean((theta-theta0)^2), xhat = open-loop Euler rollout (5 sub-steps per interval).
//       16 lanes per row: lane j owns coefficient j, its forward sensitivity d xhat / d theta_j and
//       row j of the inverse-Hessian approximation; dot products and mat-vecs go through half-warp
//       shuffles.  Line search: strong Wolfe (c1=1e-4, c2=0.9) with bracketing + zoom (Nocedal & Wright alg.
//       3.5/3.6) in two flavours (include/b200i.h): B200I_LS_JAX restates jax.scipy.optimize's semantics including
//       the way its zoom FAILS (signed bracket width <= 1e-10, 30 trials) and what a failed search leaves behind --
//       with gtol = 1e-5 (jax ignores the reference's tol=1e-12) this reproduces both INSITE lines of the reference's
//       logs (5e-15 and 2e-6 relative); B200I_LS_ROBUST accepts the best sufficient-decrease point at the FP64 noise
//       floor and never returns a point worse than theta0.
#include <stdlib.h>
#include "sim_math.cuh"
#include "stlsq.cuh"

namespace b200i {

// ------------------------------------------------------------------------------------------------
// K5b
// ------------------------------------------------------------------------------------------------
// ridge-to-prior STLSQ of one row from its per-treatment sums (see b200i_stlsq_batched in include/b200i.h)
// estimator 0: ridge shrunk to the prior on mean-normalised normal equations (north-star estimator).
// estimator 1: the ridge / threshold loop of the reference's dormant LSQIntialMask (pkpd/utils.py:244-327): the prior only
//              supplies the initial support (:251-253), the ridge pulls towards ZERO on the un-normalised normal
//              equations (sklearn ridge_regression, :228) -- i.e. what pkpd_simulation.py:795-797 keeps (unbias=False).
__device__ __forceinline__ void ridge_prior_solve(const PatientGram &pg, double u, const double *s_prior, double support_tol,
                                                  double lam, double threshold, int max_iter, double *__restrict__ out16,
                                                  int estimator = 0)
{
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double g15[B200I_GRAM_PER_TREATMENT], G[4][4], b[4], c[4], pr[4];
        expand_gram(pg.s[a], u, g15);
        unsigned ind = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            pr[j] = s_prior[a * 4 + j];
            if (fabs(pr[j]) > support_tol) ind |= 1u << j;
        }
        const double cnt = g15[14];
        if (cnt > 0.0 && ind != 0) {
            const double inv = estimator == 1 ? 1.0 : 1.0 / cnt;   // mean-normalised normal equations
#pragma unroll
            for (int j = 0; j < B200I_GRAM_PER_TREATMENT; ++j) g15[j] *= inv;
            unpack_gram(g15, G, b);
            for (int it = 0; it < max_iter; ++it) {
                if (!solve_spd4(G, b, ind, lam, estimator == 1 ? nullptr : pr, c)) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) c[j] = ((ind >> j) & 1u) ? pr[j] : 0.0;
                    break;
                }
                unsigned big = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (((ind >> j) & 1u) && fabs(c[j]) >= threshold) big |= 1u << j;
                if (big == ind) break;
                ind = big;
                if (ind == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) c[j] = 0.0;
                    break;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = pr[j];   // treatment never observed in the window: keep the prior
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) out16[a * 4 + j] = c[j];
    }
}

// The same with the treatment loop kept as a loop (one copy of the solver in the instruction stream instead of four):
// sums = the row's 4 x 5 per-treatment sums and out = its 16 coefficients, both in shared memory.
__device__ __forceinline__ void ridge_prior_solve_rolled(const double *sums, double u, const double *s_prior,
                                                         double support_tol, double lam, double threshold, int max_iter,
                                                         double *out, int estimator = 0)
{
#pragma unroll 1
    for (int a = 0; a < 4; ++a) {
        double s5[5], g15[B200I_GRAM_PER_TREATMENT], G[4][4], b[4], c[4], pr[4];
#pragma unroll
        for (int m = 0; m < 5; ++m) s5[m] = sums[a * 5 + m];
        expand_gram(s5, u, g15);
        unsigned ind = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            pr[j] = s_prior[a * 4 + j];
            if (fabs(pr[j]) > support_tol) ind |= 1u << j;
        }
        const double cnt = g15[14];
        if (cnt > 0.0 && ind != 0) {
            const double inv = estimator == 1 ? 1.0 : 1.0 / cnt;   // mean-normalised normal equations
#pragma unroll
            for (int j = 0; j < B200I_GRAM_PER_TREATMENT; ++j) g15[j] *= inv;
            unpack_gram(g15, G, b);
            for (int it = 0; it < max_iter; ++it) {
                if (!solve_spd4(G, b, ind, lam, estimator == 1 ? nullptr : pr, c)) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) c[j] = ((ind >> j) & 1u) ? pr[j] : 0.0;
                    break;
                }
                unsigned big = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (((ind >> j) & 1u) && fabs(c[j]) >= threshold) big |= 1u << j;
                if (big == ind) break;
                ind = big;
                if (ind == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) c[j] = 0.0;
                    break;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = pr[j];   // treatment never observed in the window: keep the prior
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) out[a * 4 + j] = c[j];
    }
}

__global__ void __launch_bounds__(128)
stlsq_batched_kernel(int64_t rows, int W, double fd_dt, const double *__restrict__ x, const uint8_t *__restrict__ codes,
                     const int *__restrict__ fit_len, const double *__restrict__ static_u,
                     const double *__restrict__ prior, double support_tol, double lam, double threshold, int max_iter,
                     const double *__restrict__ dts, int dts_per_row, double *__restrict__ coefs_out, int estimator)
{
    __shared__ double s_prior[16];
    if (threadIdx.x < 16) s_prior[threadIdx.x] = prior[threadIdx.x];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double *dr = dts ? dts + (dts_per_row ? r * W : 0) : nullptr;
    int n = fit_len[r];
    if (n > W - 1) n = W - 1;
    const double u = static_u[r];
    const double *xr = x + r * W;
    const uint8_t *cr = codes + r * W;
    PatientGram pg;
    pg.clear();
    if (n > 0) {
        double x0 = xr[0];
        int a0 = cr[0] & 3;
        for (int k = 0; k < n; ++k) {
            const double x1 = xr[k + 1];
            const int a1 = cr[k + 1 < W ? k + 1 : k] & 3;
            const double xdot = __ddiv_rn(__dsub_rn(x1, x0), dr ? dr[k] : fd_dt);
            pg.add(a0, x0, xdot);
            if (k == n - 1 || a1 != a0) pg.add(a0, x1, xdot);
            x0 = x1;
            a0 = a1;
        }
    }
    ridge_prior_solve(pg, u, s_prior, support_tol, lam, threshold, max_iter, coefs_out + r * 16, estimator);
}

// Tiled K5b (W <= K5_MAXW): a warp owns 32 consecutive rows; their code bytes arrive by coalesced 16-byte loads, the
// volumes (float64, or float32 storage: BASELINE config C4 "FP32 vs FP64") and -- with irregular sampling -- the
// per-row interval lengths travel 16 columns at a time through [32][17] staging tiles (row segments of 128 bytes in
// global memory, conflict free in shared memory); a thread walks one row, the warp stops at its longest fit window,
// and the 16 coefficients per row leave through the same tile as one contiguous block.  The per-thread row walk of
// stlsq_batched_kernel touches 8 of every 32 bytes it fetches and depends on L1 to see the rest (0.88 ms per 1M rows).
constexpr int K5_WARPS = 4;
constexpr int K5_CH = 16;
constexpr int K5_MAXW = 128;

template <typename X>
__global__ void __launch_bounds__(K5_WARPS * 32, 4)
stlsq_batched_tiled_kernel(int64_t rows, int W, double fd_dt, const X *__restrict__ x, const uint8_t *__restrict__ codes,
                           const int *__restrict__ fit_len, const double *__restrict__ static_u,
                           const double *__restrict__ prior, double support_tol, double lam, double threshold,
                           int max_iter, const double *__restrict__ dts, int dts_per_row, double *__restrict__ coefs_out,
                           int estimator)
{
    extern __shared__ __align__(16) uint8_t smem5[];
    __shared__ double s_prior[16];
    __shared__ double s_dtg[K5_MAXW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int code_bytes = (32 * W + 15) & ~15;
    double(*s_x)[K5_CH + 1] = reinterpret_cast<double(*)[K5_CH + 1]>(smem5) + (size_t)warp * 32;
    double(*s_t)[K5_CH + 1] = reinterpret_cast<double(*)[K5_CH + 1]>(
                                  smem5 + (size_t)K5_WARPS * 32 * (K5_CH + 1) * 8 + (size_t)K5_WARPS * 32 * 21 * 8 +
                                  (size_t)K5_WARPS * code_bytes) + (size_t)warp * 32;
    // layout: [x tiles][per-row sums][code bytes][interval-length tiles, only with per-row dts]
    double(*s_pg)[21] = reinterpret_cast<double(*)[21]>(smem5 + (size_t)K5_WARPS * 32 * (K5_CH + 1) * 8) + (size_t)warp * 32;
    uint8_t *s_code = smem5 + (size_t)K5_WARPS * 32 * (K5_CH + 1) * 8 + (size_t)K5_WARPS * 32 * 21 * 8 +
                      (size_t)warp * code_bytes;
    if (tid < 16) s_prior[tid] = prior[tid];
    if (dts && !dts_per_row)
        for (int k = tid; k < W; k += blockDim.x) s_dtg[k] = dts[k];
    __syncthreads();
    const bool codes16 = (reinterpret_cast<uintptr_t>(codes) & 15u) == 0;
    const int64_t ntiles = (rows + 31) / 32;
    for (int64_t tile = (int64_t)blockIdx.x * K5_WARPS + warp; tile < ntiles; tile += (int64_t)gridDim.x * K5_WARPS) {
        const int64_t first = tile * 32;
        const int nrows = (int)((rows - first < 32) ? (rows - first) : 32);
        const int64_t r = first + lane;
        const bool live = lane < nrows;
        int n = 0;
        double u = 0.0;
        if (live) {
            n = fit_len[r];
            if (n > W - 1) n = W - 1;
            if (n < 0) n = 0;
            u = static_u[r];
        }
        int nmax = n;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
        __syncwarp();
        if (nmax > 0) {
            const int nbytes = nrows * W;
            const uint8_t *g = codes + first * W;
            int done = 0;
            if (codes16) {
                const int n16 = nbytes >> 4;
                const uint4 *g4 = reinterpret_cast<const uint4 *>(g);
                uint4 *s4 = reinterpret_cast<uint4 *>(s_code);
                for (int e = lane; e < n16; e += 32) s4[e] = __ldg(g4 + e);
                done = n16 << 4;
            }
            for (int e = done + lane; e < nbytes; e += 32) s_code[e] = g[e];
        }
        __syncwarp();
        PatientGram pg;
        pg.clear();
        double x0 = 0.0, dt_k = fd_dt;
        const uint8_t *cr = s_code + lane * W;
        // column j of the row: j = 0 starts the walk, j >= 1 closes transition k = j - 1
        for (int j0 = 0; j0 <= nmax; j0 += K5_CH) {
            const int nc = (nmax + 1 - j0 < K5_CH) ? (nmax + 1 - j0) : K5_CH;
            const X *gx = x + first * W + j0;
            if (nc == K5_CH)
                for (int e = lane; e < nrows * K5_CH; e += 32) s_x[e >> 4][e & 15] = (double)gx[(int64_t)(e >> 4) * W + (e & 15)];
            else
                for (int e = lane; e < nrows * nc; e += 32) s_x[e / nc][e % nc] = (double)gx[(int64_t)(e / nc) * W + (e % nc)];
            if (dts && dts_per_row) {
                const double *gt = dts + first * W + j0;
                if (nc == K5_CH)
                    for (int e = lane; e < nrows * K5_CH; e += 32) s_t[e >> 4][e & 15] = gt[(int64_t)(e >> 4) * W + (e & 15)];
                else
                    for (int e = lane; e < nrows * nc; e += 32) s_t[e / nc][e % nc] = gt[(int64_t)(e / nc) * W + (e % nc)];
            }
            __syncwarp();
            if (live) {
                for (int jj = 0; jj < nc; ++jj) {
                    const int j = j0 + jj;
                    const double xv = s_x[lane][jj];
                    if (j >= 1 && j <= n) {
                        const int k = j - 1;
                        const int a0 = cr[k] & 3;
                        const int a1 = cr[k + 1 < W ? k + 1 : k] & 3;
                        const double xdot = __ddiv_rn(__dsub_rn(xv, x0), dt_k);
                        // library row of the sample (x0) and, at the end of a constant-treatment snippet, of its last
                        // point (x1, backward difference = the same slope), filed under treatment a0 in one go
                        const double e = (k == n - 1 || a1 != a0) ? 1.0 : 0.0;
                        const double cx = fma(e, xv, x0), cn = 1.0 + e;
                        const double cxx = fma(e * xv, xv, x0 * x0), cd = xdot * cn, cxd = xdot * cx;
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            if (a0 == a) {
                                pg.s[a][0] += cn; pg.s[a][1] += cx; pg.s[a][2] += cxx; pg.s[a][3] += cd; pg.s[a][4] += cxd;
                            }
                        }
                    }
                    x0 = xv;
                    if (dts) dt_k = dts_per_row ? s_t[lane][jj] : s_dtg[j];   // length of the interval that starts at column j
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int m = 0; m < 5; ++m) s_pg[lane][a * 5 + m] = pg.s[a][m];
        __syncwarp();
        if (live) ridge_prior_solve_rolled(s_pg[lane], u, s_prior, support_tol, lam, threshold, max_iter, s_x[lane], estimator);
        __syncwarp();
        double *go = coefs_out + first * 16;
        for (int e = lane; e < nrows * 16; e += 32) go[e] = s_x[e >> 4][e & 15];
        __syncwarp();
    }
}

// K5b on a compact counterfactual cohort: all rows of one (patient, t) share their fit window F[0..n_fit], so a thread
// walks one patient once and emits the T-1 fits as running sums: n_fit = t + fit_offset for
// t < executed steps, else the prior.  Same sums in the same order as stlsq_batched_kernel on the dense rows.
__global__ void __launch_bounds__(128)
stlsq_prefix_kernel(int64_t n, int T, int fit_offset, double fd_dt, const double *__restrict__ F,
                    const uint8_t *__restrict__ codes, const int *__restrict__ n_steps, const double *__restrict__ static_u,
                    const double *__restrict__ prior, double support_tol, double lam, double threshold, int max_iter,
                    double *__restrict__ coefs_out)
{
    __shared__ double s_prior[16];
    if (threadIdx.x < 16) s_prior[threadIdx.x] = prior[threadIdx.x];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int W = T - 1;
    int ns = n_steps[i];
    if (ns > W) ns = W;
    const double u = static_u[i];
    const double *xr = F + i * T;
    const uint8_t *cr = codes + i * T;
    double *out = coefs_out + i * (int64_t)W * 16;
    PatientGram pg;
    pg.clear();
    // rows whose window is empty keep the prior
    for (int t = 0; t < W; ++t)
        if (t >= ns || t + fit_offset <= 0) ridge_prior_solve(pg, u, s_prior, support_tol, lam, threshold, max_iter, out + t * 16);
    const int kmax = ns - 1 + fit_offset;     // transitions 0 .. kmax-1 are used by some row
    double x0 = xr[0];
    int a0 = ((cr[0] & 1) << 1) | ((cr[0] >> 1) & 1);
    for (int k = 0; k < kmax && k < W; ++k) {
        const double x1 = xr[k + 1];
        const int c1 = cr[k + 1 < T ? k + 1 : k];
        const int a1 = ((c1 & 1) << 1) | ((c1 >> 1) & 1);
        const double xdot = __ddiv_rn(__dsub_rn(x1, x0), fd_dt);
        pg.add(a0, x0, xdot);
        PatientGram tmp = pg;
        tmp.add(a0, x1, xdot);                // the window of n_fit = k+1 ends here: backward difference at its last point
        const int t = k + 1 - fit_offset;
        if (t >= 0 && t < ns) ridge_prior_solve(tmp, u, s_prior, support_tol, lam, threshold, max_iter, out + t * 16);
        if (a1 != a0) pg = tmp;               // a treatment change makes the end point part of every longer window
        x0 = x1;
        a0 = a1;
    }
}

// ------------------------------------------------------------------------------------------------
// K7: cooperative BFGS, 16 lanes per row
// ------------------------------------------------------------------------------------------------
constexpr int BFGS_THREADS = 128;
constexpr int K7_MINB_DEFAULT = 4;
constexpr int BFGS_GROUPS = BFGS_THREADS / 16;
constexpr int BFGS_MAXW = 80;

struct RowData {
    const double *x;       // shared memory: x[0..n_fit]
    const uint8_t *codes;  // shared memory: codes[0..n_fit-1]
    const double *dts;     // shared memory: interval lengths dts[0..n_fit-1] (irregular sampling), or nullptr
    int n_fit;
    double u, h, norm, lam;
    int substeps;
};

__device__ __forceinline__ double shfl16(double v, int src, unsigned mask, int base)
{
    return __shfl_sync(mask, v, base + src);
}
__device__ __forceinline__ double sum16(double v, unsigned mask)
{
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
__device__ __forceinline__ double max16(double v, unsigned mask)
{
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o));
    return v;
}

// objective and gradient component of this lane: f_to_min_func (sindy.py:781-794) with forward
// sensitivities through the Euler rollout (predict_with_reduced_coefs :767-778, odeint pkpd/utils.py:68-90)
template <bool DTS>
__device__ __forceinline__ void eval_objective(const RowData &d, double theta, double theta0, double mask_j, int my_a,
                                               int my_m, unsigned gmask, int gbase, double &f, double &g)
{
    const double tm = theta * mask_j;
    double v = d.x[0], s = 0.0, acc = 0.0, gacc = 0.0;
    for (int k = 0; k < d.n_fit; ++k) {
        const int a = d.codes[k] & 3;
        const double c0 = shfl16(tm, 4 * a + 0, gmask, gbase), c1 = shfl16(tm, 4 * a + 1, gmask, gbase);
        const double c2 = shfl16(tm, 4 * a + 2, gmask, gbase), c3 = shfl16(tm, 4 * a + 3, gmask, gbase);
        const double c2u = c2 * d.u, dfdv = c1 + c3 * d.u;
        const bool mine = (a == my_a);
        const double h = DTS ? d.dts[k] / d.substeps : d.h;
        for (int q = 0; q < d.substeps; ++q) {
            const double vu = v * d.u;
            const double fval = ((c0 + c1 * v) + c2u) + c3 * vu;
            const double basis = my_m == 0 ? 1.0 : (my_m == 1 ? v : (my_m == 2 ? d.u : vu));
            const double dfj = mine ? basis * mask_j : 0.0;
            s = s + h * (dfdv * s + dfj);
            v = v + h * fval;
        }
        const double r = d.x[k + 1] - v;
        acc += r * r;
        gacc += -2.0 * r * s;
    }
    const double inv = 1.0 / (double)d.n_fit;
    const double diff = theta - theta0;
    const double pen = sum16(diff * diff, gmask) * (1.0 / 16.0);
    f = acc * inv / d.norm + d.lam * pen;
    g = gacc * inv / d.norm + d.lam * 2.0 * diff * (1.0 / 16.0);
}

// The same for the joint ("one ODE") model (sindy.py:503-517): one 11-term equation over (x0 = volume, u0 = chemo
// application, u1 = radio application, u2 = static feature), library order of PolynomialLibrary(degree=2,
// interaction_only=True): 1 x0 u0 u1 u2 x0u0 x0u1 x0u2 u0u1 u0u2 u1u2.  Lane j < 11 owns coefficient j; its basis value
// is (x0 or 1) * w_j(treatment code), w_j tabulated per row for the four codes; the constant part and the slope of
// the step's right-hand side are two masked sums over the lanes.  Lanes 11..15 carry zeros.
template <bool DTS>
__device__ __forceinline__ void eval_objective_joint(const RowData &d, double theta, double theta0, double mask_j,
                                                     int j, unsigned gmask, double &f, double &g)
{
    const bool is_x = (j == 1) || (j >= 5 && j <= 7);
    double wtab[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const double ch = (double)(a & 1), ra = (double)(a >> 1), st = d.u;
        double w;
        switch (j) {
            case 0: case 1: w = 1.0; break;
            case 2: case 5: w = ch; break;
            case 3: case 6: w = ra; break;
            case 4: case 7: w = st; break;
            case 8: w = ch * ra; break;
            case 9: w = ch * st; break;
            case 10: w = ra * st; break;
            default: w = 0.0;
        }
        wtab[a] = w;
    }
    const double tm = theta * mask_j;
    double v = d.x[0], s = 0.0, acc = 0.0, gacc = 0.0;
    for (int k = 0; k < d.n_fit; ++k) {
        const int a = d.codes[k] & 3;
        const double wa = a == 0 ? wtab[0] : (a == 1 ? wtab[1] : (a == 2 ? wtab[2] : wtab[3]));
        const double tw = tm * wa;
        const double c_const = sum16(is_x ? 0.0 : tw, gmask);
        const double c_x = sum16(is_x ? tw : 0.0, gmask);
        const double wj = wa * mask_j;
        const double h = DTS ? d.dts[k] / d.substeps : d.h;
        for (int q = 0; q < d.substeps; ++q) {
            const double fval = c_const + c_x * v;
            const double dfj = is_x ? wj * v : wj;
            s = s + h * (c_x * s + dfj);
            v = v + h * fval;
        }
        const double r = d.x[k + 1] - v;
        acc += r * r;
        gacc += -2.0 * r * s;
    }
    const double inv = 1.0 / (double)d.n_fit;
    const double diff = theta - theta0;
    const double pen = sum16(diff * diff, gmask) * (1.0 / 11.0);
    f = acc * inv / d.norm + d.lam * pen;
    g = gacc * inv / d.norm + d.lam * 2.0 * diff * (1.0 / 11.0);
}

// minimiser of the cubic through (a,fa,fpa), (b,fb), (c,fc); NaN if it does not exist (scipy _cubicmin)
__device__ __forceinline__ double cubicmin(double a, double fa, double fpa, double b, double fb, double c, double fc)
{
    const double C = fpa, db = b - a, dc = c - a;
    const double denom = (db * dc) * (db * dc) * (db - dc);
    if (denom == 0.0) return nan("");
    const double t0 = fb - fa - C * db, t1 = fc - fa - C * dc;
    double A = (dc * dc * t0 - db * db * t1) / denom;
    double B = (-dc * dc * dc * t0 + db * db * db * t1) / denom;
    const double radical = B * B - 3.0 * A * C;
    if (!(radical >= 0.0) || A == 0.0) return nan("");
    return a + (-B + sqrt(radical)) / (3.0 * A);
}
__device__ __forceinline__ double quadmin(double a, double fa, double fpa, double b, double fb)
{
    const double db = b - a;
    const double B = (fb - fa - fpa * db) / (db * db);
    if (!(B > 0.0)) return nan("");
    return a - fpa / (2.0 * B);
}

// PREFIX: the rows are the (patient, t) pairs of a compact counterfactual cohort (b200i_insite_bfgs_prefix): row
// r = patient * (W-1) + t reads the patient's factual trajectory x[patient, 0..W) and its factual option codes
// (2*chemo + radio, translated to chemo + 2*radio), seq_len holds the patient's executed steps and `ph` the offset of
// the fit window: n_fit = t + ph for t < executed steps, else the row does not exist (theta0, status -2).
template <int MINB, bool JOINT = false, bool PREFIX = false, bool DTS = false>
__global__ void __launch_bounds__(BFGS_THREADS, MINB)
insite_bfgs_kernel(int64_t rows, int W, double dt, int substeps, const double *__restrict__ x,
                   const uint8_t *__restrict__ codes, const int *__restrict__ seq_len, int ph,
                   const double *__restrict__ static_u, const double *__restrict__ theta0_g, double lam, double gtol,
                   int max_iter, double *__restrict__ coefs_out, int *__restrict__ status_out,
                   double *__restrict__ fval_out, const double *__restrict__ dts = nullptr, int dts_per_row = 0,
                   int ls_mode = 0)
{
    __shared__ double s_x[BFGS_GROUPS][BFGS_MAXW + 1];
    __shared__ double s_dt[DTS ? BFGS_GROUPS : 1][BFGS_MAXW];
    __shared__ uint8_t s_c[BFGS_GROUPS][BFGS_MAXW];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 4, j = threadIdx.x & 15;
    const int gbase = lane & 16;
    const unsigned gmask = 0xFFFFu << gbase;
    const int my_a = j >> 2, my_m = j & 3;
    constexpr int NP = JOINT ? 11 : 16;                       // coefficients per row
    const double theta0 = j < NP ? theta0_g[j] : 0.0;
    const double mask_j = fabs(theta0) > 1e-3 ? 1.0 : 0.0;   // coef_sparse_mask, sindy.py:589
    auto objective = [&](const RowData &d, double th, double &f, double &g) {
        if (JOINT) eval_objective_joint<DTS>(d, th, theta0, mask_j, j, gmask, f, g);
        else eval_objective<DTS>(d, th, theta0, mask_j, my_a, my_m, gmask, gbase, f, g);
    };
    const int64_t ngroups = (int64_t)gridDim.x * BFGS_GROUPS;
    const int64_t iters = (rows + ngroups - 1) / ngroups;

    for (int64_t itr = 0; itr < iters; ++itr) {
        const int64_t r = itr * ngroups + (int64_t)blockIdx.x * BFGS_GROUPS + grp;
        const bool valid = r < rows;
        int n_fit = 0;
        int64_t src = r;
        if (valid) {
            if (PREFIX) {
                src = r / (W - 1);
                const int t = (int)(r - src * (W - 1));
                n_fit = (t < seq_len[src]) ? t + ph : 0;
            } else {
                n_fit = seq_len[r] - ph;
            }
            if (n_fit > W - 1) n_fit = W - 1;
        }
        __syncwarp(gmask);
        if (valid && n_fit > 0) {
            for (int k = j; k <= n_fit; k += 16) s_x[grp][k] = x[src * W + k];
            for (int k = j; k < n_fit; k += 16) {
                const int c = codes[src * W + k];
                s_c[grp][k] = (uint8_t)(PREFIX ? (((c & 1) << 1) | ((c >> 1) & 1)) : c);
            }
            if (DTS)
                for (int k = j; k < n_fit; k += 16) s_dt[grp][k] = dts[(dts_per_row ? src * W : 0) + k];
        }
        __syncwarp(gmask);
        if (!valid) continue;
        if (n_fit <= 0) {   // sequence_length <= projection_horizon: population coefficients (sindy.py:571-585)
            if (j < NP) coefs_out[r * NP + j] = theta0;
            if (j == 0) { status_out[r] = -2; fval_out[2 * r] = 0.0; fval_out[2 * r + 1] = 0.0; }
            continue;
        }
        RowData d;
        d.x = s_x[grp]; d.codes = s_c[grp]; d.dts = DTS ? s_dt[grp] : nullptr; d.n_fit = n_fit; d.u = static_u[src];
        d.h = dt / substeps; d.substeps = substeps; d.norm = 1.0; d.lam = lam;

        double theta = theta0, f, g;
        objective(d, theta, f, g);
        const double start_res = f;                 // norm_const = 1 (sindy.py:591-603)
        d.norm = 2.5 * start_res;                   // sindy.py:616
        int status = 0, it = 0;
        double f0 = 0.0;
        if (!(d.norm > 0.0) || !isfinite(d.norm)) {
            status = 4;                             // perfect fit already (or non-finite): keep theta0
        } else {
            objective(d, theta, f, g);
            f0 = f;
            double Hrow[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) Hrow[i] = (i == j) ? 1.0 : 0.0;
            double old_old = f + sqrt(sum16(g * g, gmask)) * 0.5;   // jax: f_0 + ||g_0|| / 2
            for (it = 0; it < max_iter; ++it) {
                if (max16(fabs(g), gmask) < gtol) { status = 0; break; }
                // p = -H g
                double p = 0.0;
#pragma unroll
                for (int i = 0; i < 16; ++i) p -= Hrow[i] * shfl16(g, i, gmask, gbase);
                double dphi0 = sum16(g * p, gmask);
                if (!(dphi0 < 0.0)) {               // not a descent direction: reset to steepest descent
#pragma unroll
                    for (int i = 0; i < 16; ++i) Hrow[i] = (i == j) ? 1.0 : 0.0;
                    p = -g;
                    dphi0 = sum16(g * p, gmask);
                    if (!(dphi0 < 0.0)) { status = 0; break; }
                }
                // ---- strong-Wolfe line search -------------------------------------------------------
                const double c1 = 1e-4, c2 = 0.9, phi0 = f;
                double a_prev = 0.0, phi_prev = phi0, dphi_prev = dphi0;
                double cand = 1.01 * 2.0 * (phi0 - old_old) / dphi0;
                double alpha = (cand > 1.0 || !(cand > 0.0)) ? 1.0 : cand;
                if (ls_mode == 1) alpha = cand > 1.0 ? 1.0 : cand;      // jax: where(candidate > 1, 1.0, candidate)
                double a_star = 0.0, f_star = f, g_star = g;
                bool found = false, ls_failed = false, zoom_failed = false;
                double lo = 0, hi = 0, phi_lo = 0, dphi_lo = 0, phi_hi = 0;
                bool need_zoom = false;
                for (int i = 1; i <= 10; ++i) {
                    double fi, gi;
                    objective(d, theta + alpha * p, fi, gi);
                    const double dphi_i = sum16(gi * p, gmask);
                    if (!isfinite(fi) || fi > phi0 + c1 * alpha * dphi0 || (fi >= phi_prev && i > 1)) {
                        lo = a_prev; phi_lo = phi_prev; dphi_lo = dphi_prev; hi = alpha; phi_hi = fi;
                        need_zoom = true; break;
                    }
                    if (fabs(dphi_i) <= -c2 * dphi0) { a_star = alpha; f_star = fi; g_star = gi; found = true; break; }
                    if (dphi_i >= 0.0) {
                        lo = alpha; phi_lo = fi; dphi_lo = dphi_i; hi = a_prev; phi_hi = phi_prev;
                        need_zoom = true; break;
                    }
                    a_prev = alpha; phi_prev = fi; dphi_prev = dphi_i;
                    alpha *= 2.0;
                    if (i == 10) ls_failed = true;
                }
                if (need_zoom) {
                    double a_rec = 0.0, phi_rec = phi0;
                    bool have_rec = false;
                    ls_failed = true;
                    for (int z = 0; z < 30; ++z) {
                        const double dalpha = hi - lo;
                        // ls_mode 1: jax's _zoom declares failure when the SIGNED bracket width a_hi - a_lo is <= 1e-10
                        // (float64), i.e. also at once for a reversed bracket (a_hi < a_lo); the iteration that notices
                        // it still evaluates its trial point
                        if (ls_mode == 1 && dalpha <= 1e-10) zoom_failed = true;
                        const double a_min = dalpha < 0 ? hi : lo, a_max = dalpha < 0 ? lo : hi;
                        double aj = nan("");
                        if (have_rec) {
                            const double cchk = 0.2 * dalpha;
                            aj = cubicmin(lo, phi_lo, dphi_lo, hi, phi_hi, a_rec, phi_rec);
                            if (ls_mode ? (isnan(aj) || !(aj > a_min + cchk) || !(aj < a_max - cchk))
                                        : (isnan(aj) || aj > a_max - fabs(cchk) || aj < a_min + fabs(cchk))) aj = nan("");
                        }
                        if (isnan(aj)) {
                            const double qchk = 0.1 * dalpha;
                            aj = quadmin(lo, phi_lo, dphi_lo, hi, phi_hi);
                            if (ls_mode ? (isnan(aj) || !(aj > a_min + qchk) || !(aj < a_max - qchk))
                                        : (isnan(aj) || aj > a_max - fabs(qchk) || aj < a_min + fabs(qchk))) aj = lo + 0.5 * dalpha;
                        }
                        double fj, gj;
                        objective(d, theta + aj * p, fj, gj);
                        const double dphi_j = sum16(gj * p, gmask);
                        if (!isfinite(fj) || fj > phi0 + c1 * aj * dphi0 || fj >= phi_lo) {
                            a_rec = hi; phi_rec = phi_hi; have_rec = true;
                            hi = aj; phi_hi = fj;
                        } else {
                            if (fabs(dphi_j) <= -c2 * dphi0) {
                                a_star = aj; f_star = fj; g_star = gj; found = true; ls_failed = false; break;
                            }
                            if (dphi_j * (hi - lo) >= 0.0) {
                                a_rec = hi; phi_rec = phi_hi; hi = lo; phi_hi = phi_lo;
                                if (ls_mode) { a_rec = lo; phi_rec = phi_lo; }   // jax applies lo_to_j after hi_to_lo
                            } else {
                                a_rec = lo; phi_rec = phi_lo;
                            }
                            have_rec = true;
                            lo = aj; phi_lo = fj; dphi_lo = dphi_j;
                            // remember the best sufficient-decrease point in case the curvature test never passes
                            a_star = aj; f_star = fj; g_star = gj;
                        }
                        if (zoom_failed) break;
                        if (ls_mode == 0 && fabs(hi - lo) <= 1e-16 * fmax(1.0, fabs(lo))) break;
                    }
                }
                if (ls_mode == 1 && need_zoom && (zoom_failed || !found)) {
                    // jax: a failed zoom ends BFGS (status 3 = 2 + line-search status 1).  The state it leaves behind is
                    // x + a p with a = the trial point if that happened to satisfy both Wolfe conditions, else the
                    // initial a_star = 1 of _ZoomState; the reference (sindy.py:628-631) then discards it.
                    const double a_left = found ? a_star : 1.0;
                    theta += a_left * p;
                    objective(d, theta, f, g);
                    status = 3;
                    ++it;                 // jax counts the iteration whose line search failed
                    break;
                }
                if (!found) {
                    // line search exhausted (typically at the FP64 noise floor of the objective): accept the
                    // best sufficient-decrease point if it improves f, then stop
                    if (a_star > 0.0 && f_star < f) { theta += a_star * p; f = f_star; g = g_star; }
                    status = (ls_mode == 1) ? 5 : (ls_failed ? 3 : 5);   // mode 1 gets here only from the bracketing phase
                    if (ls_mode == 1) ++it;
                    break;
                }
                // ---- BFGS update of the inverse Hessian ---------------------------------------------
                const double s_k = a_star * p, y_k = g_star - g;
                const double f_old = f;
                theta += s_k; old_old = f; f = f_star; g = g_star;
                const double sy = sum16(s_k * y_k, gmask);
                double rho = 1.0 / sy;
                const bool keep_H = ls_mode == 1 && !isfinite(rho);   // jax: H_kp1 = where(isfinite(rho_k), H_kp1, H_k)
                if (!isfinite(rho)) rho = keep_H ? 0.0 : 1000.0;
                double Hy = 0.0;
#pragma unroll
                for (int i = 0; i < 16; ++i) Hy += Hrow[i] * shfl16(y_k, i, gmask, gbase);
                const double yHy = sum16(y_k * Hy, gmask);
                const double coef = rho * rho * yHy + rho;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const double s_i = shfl16(s_k, i, gmask, gbase), Hy_i = shfl16(Hy, i, gmask, gbase);
                    Hrow[i] = Hrow[i] - rho * (s_k * Hy_i + Hy * s_i) + coef * s_k * s_i;
                }
                if (ls_mode == 0 && fabs(f_old - f) <= 1e-15 * fmax(fabs(f), 1e-300)) { status = 0; ++it; break; }
            }
            if (it >= max_iter && status == 0) status = 1;
            if (ls_mode == 0 && (!(f <= f0) || !isfinite(f))) { theta = theta0; f = f0; status = 6; }   // never accept a worse point
        }
        if (j < NP) coefs_out[r * NP + j] = theta;
        if (j == 0) { status_out[r] = status | (it << 8); fval_out[2 * r] = f0; fval_out[2 * r + 1] = f; }
    }
}

}  // namespace b200i

using namespace b200i;

template <typename X>
static int stlsq_batched_impl(int64_t rows, int32_t W, double fd_dt, const X *x, const uint8_t *codes,
                              const int32_t *fit_len, const double *static_feature, const double *prior,
                              double support_tol, double lam, double threshold, int32_t max_iter, const double *dts,
                              int32_t dts_per_row, double *coefs_out, void *stream, int32_t estimator = 0)
{
    B200I_REQUIRE(rows >= 0, B200I_E_ARG, "stlsq_batched: negative rows");
    B200I_REQUIRE(estimator == 0 || estimator == 1, B200I_E_ARG, "stlsq_batched: estimator %d (0 ridge-to-prior, 1 LSQIntialMask)", estimator);
    if (rows == 0) return 0;
    B200I_REQUIRE(x && codes && fit_len && static_feature && prior && coefs_out, B200I_E_ARG, "stlsq_batched: NULL argument");
    B200I_REQUIRE(W >= 2 && (dts != nullptr || fd_dt > 0) && lam > 0 && threshold >= 0 && max_iter >= 1, B200I_E_ARG,
                  "stlsq_batched: need W >= 2, fd_dt > 0, lam > 0 (the per-row design is rank deficient), threshold >= 0");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (W <= K5_MAXW) {
        const int code_bytes = (32 * W + 15) & ~15;
        const int smem = K5_WARPS * (32 * (K5_CH + 1) * 8 + 32 * 21 * 8 + code_bytes +
                                     ((dts && dts_per_row) ? 32 * (K5_CH + 1) * 8 : 0));
        const void *kern = reinterpret_cast<const void *>(stlsq_batched_tiled_kernel<X>);
        int per_sm = 1;
        {
            int rc0 = ensure_dyn_smem(kern, smem, K5_WARPS * 32, &per_sm);
            if (rc0) return rc0;
        }
        if (per_sm < 1) per_sm = 1;
        const int64_t ntiles = (rows + 31) / 32;
        int64_t grid = (ntiles + K5_WARPS - 1) / K5_WARPS;
        const int64_t cap = (int64_t)num_sms() * per_sm;
        if (grid > cap) grid = cap;
        stlsq_batched_tiled_kernel<X><<<(unsigned)grid, K5_WARPS * 32, smem, st>>>(
            rows, W, fd_dt, x, codes, fit_len, static_feature, prior, support_tol, lam, threshold, max_iter, dts, dts_per_row,
            coefs_out, estimator);
        return check_cuda(cudaGetLastError(), "stlsq_batched launch");
    }
    B200I_REQUIRE(sizeof(X) == sizeof(double), B200I_E_UNSUPPORTED, "stlsq_batched_f32: W=%d > %d", W, K5_MAXW);
    const unsigned grid = (unsigned)((rows + 127) / 128);
    stlsq_batched_kernel<<<grid, 128, 0, st>>>(rows, W, fd_dt, reinterpret_cast<const double *>(x), codes, fit_len,
                                               static_feature, prior, support_tol, lam, threshold, max_iter, dts, dts_per_row,
                                               coefs_out, estimator);
    return check_cuda(cudaGetLastError(), "stlsq_batched launch");
}

extern "C" int b200i_stlsq_batched(int64_t rows, int32_t W, double fd_dt, const double *x, const uint8_t *codes,
                                   const int32_t *fit_len, const double *static_feature, const double *prior,
                                   double support_tol, double lam, double threshold, int32_t max_iter,
                                   double *coefs_out, void *stream)
{
    return stlsq_batched_impl<double>(rows, W, fd_dt, x, codes, fit_len, static_feature, prior, support_tol, lam, threshold,
                                      max_iter, nullptr, 0, coefs_out, stream);
}

extern "C" int b200i_stlsq_batched_dts(int64_t rows, int32_t W, const double *x, const float *x_f32, const uint8_t *codes,
                                       const int32_t *fit_len, const double *static_feature, const double *prior,
                                       double support_tol, double lam, double threshold, int32_t max_iter,
                                       double fd_dt, const double *dts, int32_t dts_per_row, int32_t estimator,
                                       double *coefs_out, void *stream)
{
    B200I_REQUIRE((x != nullptr) != (x_f32 != nullptr), B200I_E_ARG, "stlsq_batched_dts: pass exactly one of x / x_f32");
    if (x_f32)
        return stlsq_batched_impl<float>(rows, W, fd_dt, x_f32, codes, fit_len, static_feature, prior, support_tol, lam,
                                         threshold, max_iter, dts, dts_per_row, coefs_out, stream, estimator);
    return stlsq_batched_impl<double>(rows, W, fd_dt, x, codes, fit_len, static_feature, prior, support_tol, lam, threshold,
                                      max_iter, dts, dts_per_row, coefs_out, stream, estimator);
}

static int insite_bfgs_impl(int64_t rows, int32_t W, double dt, int32_t substeps, const double *x, const uint8_t *codes,
                            const int32_t *sequence_lengths, int32_t projection_horizon, const double *static_feature,
                            const double *theta0, double lam, double gtol, int32_t max_iter, int32_t line_search,
                            const double *dts, int32_t dts_per_row, double *coefs_out, int32_t *status_out, double *fval_out,
                            void *stream)
{
    B200I_REQUIRE(rows >= 0, B200I_E_ARG, "insite_bfgs: negative rows");
    B200I_REQUIRE(line_search == B200I_LS_JAX || line_search == B200I_LS_ROBUST, B200I_E_ARG, "insite_bfgs: line_search %d", line_search);
    if (rows == 0) return 0;
    B200I_REQUIRE(x && codes && sequence_lengths && static_feature && theta0 && coefs_out && status_out && fval_out,
                  B200I_E_ARG, "insite_bfgs: NULL argument");
    B200I_REQUIRE(W >= 2 && W <= BFGS_MAXW, B200I_E_UNSUPPORTED, "insite_bfgs: W=%d outside [2,%d]", W, BFGS_MAXW);
    B200I_REQUIRE((dts != nullptr || dt > 0) && substeps >= 1 && lam >= 0 && max_iter >= 1, B200I_E_ARG,
                  "insite_bfgs: bad scalar argument");
    int64_t grid = (rows + BFGS_GROUPS - 1) / BFGS_GROUPS;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (grid > cap) grid = cap;
    // register budget for 4 CTAs per SM (measured: 3 -> 85 ms, 4 -> 69 ms, 5 -> 76 ms per 200k rows)
    if (dts)
        insite_bfgs_kernel<K7_MINB_DEFAULT, false, false, true><<<(unsigned)grid, BFGS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
            rows, W, dt, substeps, x, codes, sequence_lengths, projection_horizon, static_feature, theta0, lam, gtol,
            max_iter, coefs_out, status_out, fval_out, dts, dts_per_row, line_search);
    else
        insite_bfgs_kernel<K7_MINB_DEFAULT, false><<<(unsigned)grid, BFGS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
            rows, W, dt, substeps, x, codes, sequence_lengths, projection_horizon, static_feature, theta0, lam, gtol,
            max_iter, coefs_out, status_out, fval_out, nullptr, 0, line_search);
    return check_cuda(cudaGetLastError(), "insite_bfgs launch");
}

extern "C" int b200i_insite_bfgs(int64_t rows, int32_t W, double dt, int32_t substeps, const double *x,
                                 const uint8_t *codes, const int32_t *sequence_lengths, int32_t projection_horizon,
                                 const double *static_feature, const double *theta0, double lam, double gtol,
                                 int32_t max_iter, int32_t line_search, double *coefs_out, int32_t *status_out,
                                 double *fval_out, void *stream)
{
    return insite_bfgs_impl(rows, W, dt, substeps, x, codes, sequence_lengths, projection_horizon, static_feature, theta0,
                            lam, gtol, max_iter, line_search, nullptr, 0, coefs_out, status_out, fval_out, stream);
}

extern "C" int b200i_insite_bfgs_dts(int64_t rows, int32_t W, int32_t substeps, const double *x, const uint8_t *codes,
                                     const int32_t *sequence_lengths, int32_t projection_horizon,
                                     const double *static_feature, const double *theta0, double lam, double gtol,
                                     int32_t max_iter, int32_t line_search, const double *dts, int32_t dts_per_row,
                                     double *coefs_out, int32_t *status_out, double *fval_out, void *stream)
{
    B200I_REQUIRE(dts != nullptr, B200I_E_ARG, "insite_bfgs_dts: dts is NULL");
    return insite_bfgs_impl(rows, W, 0.0, substeps, x, codes, sequence_lengths, projection_horizon, static_feature, theta0,
                            lam, gtol, max_iter, line_search, dts, dts_per_row, coefs_out, status_out, fval_out, stream);
}

extern "C" int b200i_insite_bfgs_joint(int64_t rows, int32_t W, double dt, int32_t substeps, const double *x,
                                       const uint8_t *codes, const int32_t *sequence_lengths,
                                       int32_t projection_horizon, const double *static_feature, const double *theta0,
                                       double lam, double gtol, int32_t max_iter, int32_t line_search, double *coefs_out,
                                       int32_t *status_out, double *fval_out, void *stream)
{
    B200I_REQUIRE(rows >= 0, B200I_E_ARG, "insite_bfgs_joint: negative rows");
    B200I_REQUIRE(line_search == B200I_LS_JAX || line_search == B200I_LS_ROBUST, B200I_E_ARG, "insite_bfgs_joint: line_search %d", line_search);
    if (rows == 0) return 0;
    B200I_REQUIRE(x && codes && sequence_lengths && static_feature && theta0 && coefs_out && status_out && fval_out,
                  B200I_E_ARG, "insite_bfgs_joint: NULL argument");
    B200I_REQUIRE(W >= 2 && W <= BFGS_MAXW, B200I_E_UNSUPPORTED, "insite_bfgs_joint: W=%d outside [2,%d]", W, BFGS_MAXW);
    B200I_REQUIRE(dt > 0 && substeps >= 1 && lam >= 0 && max_iter >= 1, B200I_E_ARG,
                  "insite_bfgs_joint: bad scalar argument");
    int64_t grid = (rows + BFGS_GROUPS - 1) / BFGS_GROUPS;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (grid > cap) grid = cap;
    insite_bfgs_kernel<K7_MINB_DEFAULT, true><<<(unsigned)grid, BFGS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        rows, W, dt, substeps, x, codes, sequence_lengths, projection_horizon, static_feature, theta0, lam, gtol,
        max_iter, coefs_out, status_out, fval_out, nullptr, 0, line_search);
    return check_cuda(cudaGetLastError(), "insite_bfgs_joint launch");
}

extern "C" int b200i_insite_bfgs_prefix(int64_t n, int32_t T, int32_t fit_offset, double dt, int32_t substeps,
                                        const double *factual, const uint8_t *codes, const int32_t *n_steps,
                                        const double *static_feature, const double *theta0, double lam, double gtol,
                                        int32_t max_iter, int32_t line_search, double *coefs_out, int32_t *status_out,
                                        double *fval_out, void *stream)
{
    B200I_REQUIRE(n >= 0, B200I_E_ARG, "insite_bfgs_prefix: negative n");
    B200I_REQUIRE(line_search == B200I_LS_JAX || line_search == B200I_LS_ROBUST, B200I_E_ARG, "insite_bfgs_prefix: line_search %d", line_search);
    if (n == 0) return 0;
    B200I_REQUIRE(factual && codes && n_steps && static_feature && theta0 && coefs_out && status_out && fval_out,
                  B200I_E_ARG, "insite_bfgs_prefix: NULL argument");
    B200I_REQUIRE(T >= 3 && T <= BFGS_MAXW, B200I_E_UNSUPPORTED, "insite_bfgs_prefix: T=%d outside [3,%d]", T, BFGS_MAXW);
    B200I_REQUIRE(fit_offset >= 0 && fit_offset <= 1, B200I_E_ARG,
                  "insite_bfgs_prefix: fit_offset is 0 (one-step rows: t transitions) or 1 (sequence rows: t+1)");
    B200I_REQUIRE(dt > 0 && substeps >= 1 && lam >= 0 && max_iter >= 1, B200I_E_ARG, "insite_bfgs_prefix: bad scalar argument");
    const int64_t rows = n * (T - 1);
    int64_t grid = (rows + BFGS_GROUPS - 1) / BFGS_GROUPS;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (grid > cap) grid = cap;
    insite_bfgs_kernel<K7_MINB_DEFAULT, false, true><<<(unsigned)grid, BFGS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        rows, T, dt, substeps, factual, codes, n_steps, fit_offset, static_feature, theta0, lam, gtol, max_iter, coefs_out,
        status_out, fval_out, nullptr, 0, line_search);
    return check_cuda(cudaGetLastError(), "insite_bfgs_prefix launch");
}

extern "C" int b200i_stlsq_prefix(int64_t n, int32_t T, int32_t fit_offset, double fd_dt, const double *factual,
                                  const uint8_t *codes, const int32_t *n_steps, const double *static_feature,
                                  const double *prior, double support_tol, double lam, double threshold, int32_t max_iter,
                                  double *coefs_out, void *stream)
{
    B200I_REQUIRE(n >= 0, B200I_E_ARG, "stlsq_prefix: negative n");
    if (n == 0) return 0;
    B200I_REQUIRE(factual && codes && n_steps && static_feature && prior && coefs_out, B200I_E_ARG, "stlsq_prefix: NULL argument");
    B200I_REQUIRE(T >= 3 && fd_dt > 0 && lam > 0 && threshold >= 0 && max_iter >= 1 && fit_offset >= 0 && fit_offset <= 1,
                  B200I_E_ARG, "stlsq_prefix: need T >= 3, fd_dt > 0, lam > 0, threshold >= 0, fit_offset in {0,1}");
    const unsigned grid = (unsigned)((n + 127) / 128);
    stlsq_prefix_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(n, T, fit_offset, fd_dt, factual, codes, n_steps,
                                                                             static_feature, prior, support_tol, lam,
                                                                             threshold, max_iter, coefs_out);
    return check_cuda(cudaGetLastError(), "stlsq_prefix launch");
}
