// fit_rollout.cu -- K5 (population STLSQ), K6 (closed-form ODE rollout), treatment codes and the
// masked squared-error reductions of the RMSE metrics.
#include "stats_reduce.cuh"
#include "stlsq.cuh"

namespace b200i {

// ---- K5 ------------------------------------------------------------------------------------------
__global__ void stlsq_population_kernel(const double *__restrict__ stats, double threshold, double alpha, int max_iter,
                                        double *__restrict__ coefs, int *__restrict__ support)
{
    const int a = threadIdx.x;
    if (a >= 4) return;
    double G[4][4], b[4], c[4];
    unpack_gram(stats + a * B200I_GRAM_PER_TREATMENT, G, b);
    unsigned ind = 0;
    if (stats[a * B200I_GRAM_PER_TREATMENT + 14] > 0.0)
        ind = stlsq4(G, b, threshold, alpha, max_iter, 0xFu, c);
    else
        for (int j = 0; j < 4; ++j) c[j] = 0.0;
    for (int j = 0; j < 4; ++j) {
        coefs[a * 4 + j] = c[j];
        support[a * 4 + j] = (int)((ind >> j) & 1u);
    }
}

// ---- K5j: joint model (one ODE over [x0, chemo, radio, static]; sindy.py:185-204 with joint_model=True) ------------
// The 11-term library [1,x0,u0,u1,u2,x0u0,x0u1,x0u2,u0u1,u0u2,u1u2] (PolynomialLibrary(degree=2,
// interaction_only=True) on four inputs) evaluated at a treatment (u0,u1) = (chemo,radio) in {0,1}^2 is a linear
// image of psi = [1,x0,u2,x0u2], so its normal equations are assembled from the same per-treatment statistics that
// theta_gram (mode 1) reduces: G11 = sum_a M_a G_a M_a^T, b11 = sum_a M_a b_a.  Then STLSQ + unbias as in stlsq.cuh,
// for JP features (single thread, local arrays).
constexpr int JP = 11;

__device__ bool solve_spd_jp(const double (&G)[JP][JP], const double (&b)[JP], unsigned mask, double ridge, double (&c)[JP])
{
    double A[JP][JP], r[JP], sc[JP], y[JP];
    bool ok = true;
    for (int i = 0; i < JP; ++i) {
        const bool si = (mask >> i) & 1u;
        const double d = si ? G[i][i] + ridge : 1.0;
        ok = ok && (d > 0.0);
        sc[i] = si ? 1.0 / sqrt(d > 0.0 ? d : 1.0) : 1.0;
        r[i] = si ? b[i] * sc[i] : 0.0;
    }
    for (int i = 0; i < JP; ++i)
        for (int j = 0; j < JP; ++j) {
            const bool sel = ((mask >> i) & 1u) && ((mask >> j) & 1u);
            A[i][j] = sel ? (G[i][j] + ((i == j) ? ridge : 0.0)) * sc[i] * sc[j] : ((i == j) ? 1.0 : 0.0);
        }
    for (int j = 0; j < JP; ++j) {
        double d = A[j][j];
        for (int k = 0; k < j; ++k) d -= A[j][k] * A[j][k];
        ok = ok && (d > 0.0);
        d = sqrt(d > 0.0 ? d : 1.0);
        A[j][j] = d;
        for (int i = j + 1; i < JP; ++i) {
            double v = A[i][j];
            for (int k = 0; k < j; ++k) v -= A[i][k] * A[j][k];
            A[i][j] = v / d;
        }
    }
    for (int i = 0; i < JP; ++i) {
        double v = r[i];
        for (int k = 0; k < i; ++k) v -= A[i][k] * y[k];
        y[i] = v / A[i][i];
    }
    for (int i = JP - 1; i >= 0; --i) {
        double v = y[i];
        for (int k = i + 1; k < JP; ++k) v -= A[k][i] * y[k];
        y[i] = v / A[i][i];
    }
    for (int i = 0; i < JP; ++i) c[i] = ((mask >> i) & 1u) ? y[i] * sc[i] : 0.0;
    return ok;
}

__global__ void stlsq_joint_kernel(const double *__restrict__ stats, double threshold, double alpha, int max_iter,
                                   double drop_below, double *__restrict__ coefs11, int *__restrict__ support11,
                                   double *__restrict__ coefs44)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // feature f = mult_f(u0,u1) * psi[idx_f]
    const int idx[JP] = {0, 1, 0, 0, 2, 1, 1, 3, 0, 2, 2};
    double G[JP][JP], b[JP], coef[JP];
    for (int i = 0; i < JP; ++i) {
        b[i] = 0.0; coef[i] = 0.0;
        for (int j = 0; j < JP; ++j) G[i][j] = 0.0;
    }
    double total = 0.0;
    for (int a = 0; a < 4; ++a) {
        double Ga[4][4], ba[4];
        unpack_gram(stats + a * B200I_GRAM_PER_TREATMENT, Ga, ba);
        total += stats[a * B200I_GRAM_PER_TREATMENT + 14];
        const double u0 = (double)(a & 1), u1 = (double)(a >> 1);
        const double m[JP] = {1.0, 1.0, u0, u1, 1.0, u0, u1, 1.0, u0 * u1, u0, u1};
        for (int i = 0; i < JP; ++i) {
            b[i] += m[i] * ba[idx[i]];
            for (int j = 0; j < JP; ++j) G[i][j] += m[i] * m[j] * Ga[idx[i]][idx[j]];
        }
    }
    unsigned ind = (1u << JP) - 1u;
    if (total > 0.0) {
        const int n_selected0 = JP;
        unsigned prev_pattern = ind;   // history_[0] is the dense OLS initial guess
        for (int it = 0; it < max_iter; ++it) {
            if (ind == 0u) {
                for (int j = 0; j < JP; ++j) coef[j] = 0.0;
                break;
            }
            double c[JP];
            if (!solve_spd_jp(G, b, ind, alpha, c)) break;
            unsigned big = 0u;
            for (int j = 0; j < JP; ++j) {
                if (((ind >> j) & 1u) && fabs(c[j]) >= threshold) big |= 1u << j;
                else c[j] = 0.0;
                coef[j] = c[j];
            }
            ind = big;
            unsigned pattern = 0u;
            for (int j = 0; j < JP; ++j) pattern |= (coef[j] != 0.0 ? 1u : 0u) << j;
            const bool no_change = (pattern == prev_pattern);
            prev_pattern = pattern;
            if (__popc(ind) == n_selected0 || no_change) break;
        }
        if (ind != 0u) {   // unbias: ordinary least squares on the support
            double c[JP];
            if (solve_spd_jp(G, b, ind, 0.0, c))
                for (int j = 0; j < JP; ++j) coef[j] = c[j];
        }
    } else {
        ind = 0u;
    }
    for (int j = 0; j < JP; ++j) {
        coefs11[j] = coef[j];
        support11[j] = (int)((ind >> j) & 1u);
    }
    // the expression the reference integrates keeps the terms with |c| > drop_below (pkpd/utils.py:387-391);
    // restricted to a treatment it is the 4-term ODE over [1, x0, u2, x0 u2] that ode_rollout takes
    double e[JP];
    for (int j = 0; j < JP; ++j) e[j] = (fabs(coef[j]) > drop_below) ? coef[j] : 0.0;
    for (int a = 0; a < 4; ++a) {
        const double u0 = (double)(a & 1), u1 = (double)(a >> 1);
        coefs44[a * 4 + 0] = e[0] + e[2] * u0 + e[3] * u1 + e[8] * u0 * u1;
        coefs44[a * 4 + 1] = e[1] + e[5] * u0 + e[6] * u1;
        coefs44[a * 4 + 2] = e[4] + e[9] * u0 + e[10] * u1;
        coefs44[a * 4 + 3] = e[7];
    }
}

// ---- K6 ------------------------------------------------------------------------------------------
// single-rounding arithmetic in the rollout's compute type (no FMA contraction: mirrors the reference's
// separate multiply / add, pkpd/utils.py:68-71)
__device__ __forceinline__ double r_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double r_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float r_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float r_mul(float a, float b) { return __fmul_rn(a, b); }

// one interval: `substeps` explicit Euler steps  y + (c0*1 + c1*y + c2*u + c3*(y*u)) * h   (pkpd/utils.py:68-71,
// sindy.py:491-496) with the ODE of treatment code a (argmax of the one-hot, sindy.py:310 / :499)
template <typename R>
__device__ __forceinline__ R euler_interval(R v, const R (&c)[4][4], int a, R u, R h, int substeps)
{
    const R c0 = a == 0 ? c[0][0] : a == 1 ? c[1][0] : a == 2 ? c[2][0] : c[3][0];
    const R c1 = a == 0 ? c[0][1] : a == 1 ? c[1][1] : a == 2 ? c[2][1] : c[3][1];
    const R c2 = a == 0 ? c[0][2] : a == 1 ? c[1][2] : a == 2 ? c[2][2] : c[3][2];
    const R c3 = a == 0 ? c[0][3] : a == 1 ? c[1][3] : a == 2 ? c[2][3] : c[3][3];
    const R c2u = r_mul(c2, u);
    for (int sidx = 0; sidx < substeps; ++sidx) {
        R f = r_add(c0, r_mul(c1, v));
        f = r_add(f, c2u);
        f = r_add(f, r_mul(c3, r_mul(v, u)));
        v = r_add(v, r_mul(f, h));
    }
    return v;
}

// R = double: the reference's arithmetic (jax_enable_x64).  R = float: the FP32 variant of BASELINE config C4
// (same inputs and outputs in float64, state / coefficients / Euler steps in float32; agrees to ~1e-5 relative
// over 295 sub-steps, tolerance 1e-4).
//
// Tiled kernel (W <= R6_MAXW): a warp owns 32 consecutive rows and works on its own; their code bytes and per-row
// coefficient matrices are contiguous in global memory and arrive by coalesced 16-byte / 8-byte loads into shared
// memory; a thread integrates one row and parks 16 predictions at a time in a [32][17] staging tile (conflict free
// both ways) that the warp writes back as 128-byte row segments; per-row interval lengths (irregular sampling)
// travel through the same tile in the opposite direction.  ~10 KB of shared memory per warp instead of the 60 KB per
// 128 rows of the first version (3 CTAs / SM, byte-wise strided code loads: 0.61 ms per 1M x 59 rows).
constexpr int R6_WARPS = 4;
constexpr int R6_CH = 16;
constexpr int R6_MAXW = 128;

// SUB > 0: compile-time number of sub-steps (fully unrolled); SUB == 0: run-time `substeps`
template <typename R, int SUB>
__device__ __forceinline__ R euler_steps(R v, R c0, R c1, R c2u, R c3, R u, R h, int substeps)
{
    if (SUB > 0) {
#pragma unroll
        for (int sidx = 0; sidx < SUB; ++sidx) {
            R f = r_add(c0, r_mul(c1, v));
            f = r_add(f, c2u);
            f = r_add(f, r_mul(c3, r_mul(v, u)));
            v = r_add(v, r_mul(f, h));
        }
    } else {
        for (int sidx = 0; sidx < substeps; ++sidx) {
            R f = r_add(c0, r_mul(c1, v));
            f = r_add(f, c2u);
            f = r_add(f, r_mul(c3, r_mul(v, u)));
            v = r_add(v, r_mul(f, h));
        }
    }
    return v;
}

template <typename R, int SUB>
__global__ void __launch_bounds__(R6_WARPS * 32)
ode_rollout_tiled_kernel(int64_t rows, int W, double dt, int substeps, const double *__restrict__ x0,
                         const double *__restrict__ static_feature, const uint8_t *__restrict__ codes,
                         const double *__restrict__ coefs, int per_row, double drop_below,
                         const double *__restrict__ dts, int dts_per_row, double *__restrict__ pred)
{
    extern __shared__ __align__(16) uint8_t smem6[];
    __shared__ R s_coef[16];
    __shared__ double s_dt[R6_MAXW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int code_bytes = (32 * W + 15) & ~15;
    double(*s_io)[R6_CH + 1] = reinterpret_cast<double(*)[R6_CH + 1]>(smem6) + (size_t)warp * 32;
    // per-row coefficient table in the compute type, terms with |c| <= drop_below already dropped; pitch 17: the
    // 4-element row of (lane, code) sits in different banks for different lanes
    R(*s_cf)[17] = reinterpret_cast<R(*)[17]>(smem6 + (size_t)R6_WARPS * 32 * (R6_CH + 1) * 8) + (size_t)warp * 32;
    uint8_t *s_code = smem6 + (size_t)R6_WARPS * 32 * (R6_CH + 1) * 8 + (per_row ? (size_t)R6_WARPS * 32 * 17 * 8 : 0) +
                      (size_t)warp * code_bytes;
    if (!per_row && tid < 16) {
        const double c = coefs[tid];
        s_coef[tid] = (R)((fabs(c) > drop_below) ? c : 0.0);
    }
    if (dts && !dts_per_row)
        for (int k = tid; k < W; k += blockDim.x) s_dt[k] = dts[k];
    __syncthreads();
    const R h_uniform = (R)(dt / substeps);
    const bool codes16 = (reinterpret_cast<uintptr_t>(codes) & 15u) == 0;
    const int64_t ntiles = (rows + 31) / 32;
    for (int64_t tile = (int64_t)blockIdx.x * R6_WARPS + warp; tile < ntiles; tile += (int64_t)gridDim.x * R6_WARPS) {
        const int64_t first = tile * 32;
        const int nrows = (int)((rows - first < 32) ? (rows - first) : 32);
        const int64_t r = first + lane;
        const bool live = lane < nrows;
        __syncwarp();
        {   // code bytes of the tile: contiguous nrows * W bytes starting at a multiple of 32
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
        if (per_row) {
            const double *g = coefs + first * 16;
            for (int e = lane; e < nrows * 16; e += 32) {
                const double cv = g[e];
                s_cf[e >> 4][e & 15] = (R)((fabs(cv) > drop_below) ? cv : 0.0);
            }
        }
        R v = (R)0, u = (R)0;
        if (live) {
            v = (R)x0[r];
            u = (R)static_feature[r];
        }
        __syncwarp();
        const R *ctab = per_row ? &s_cf[lane][0] : &s_coef[0];
        const uint8_t *crow = s_code + (live ? lane : 0) * W;     // idle lanes walk row 0 (their results are never stored)
        for (int k0 = 0; k0 < W; k0 += R6_CH) {
            const int nc = (W - k0 < R6_CH) ? (W - k0) : R6_CH;
            if (dts && dts_per_row) {   // interval lengths of this chunk, (nrows, nc) row segments
                const double *g = dts + first * W + k0;
                if (nc == R6_CH)
                    for (int e = lane; e < nrows * R6_CH; e += 32) s_io[e >> 4][e & 15] = g[(int64_t)(e >> 4) * W + (e & 15)];
                else
                    for (int e = lane; e < nrows * nc; e += 32) s_io[e / nc][e % nc] = g[(int64_t)(e / nc) * W + (e % nc)];
                __syncwarp();
            }
            const uint8_t *cr = crow + k0;
#pragma unroll 4
            for (int kk = 0; kk < nc; ++kk) {
                const R *c = ctab + 4 * (cr[kk] & 3);   // the treatment's ODE (argmax of the one-hot, sindy.py:310 / :499)
                const R c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
                R h = h_uniform;
                if (dts) h = (R)((dts_per_row ? s_io[lane][kk] : s_dt[k0 + kk]) / substeps);
                v = euler_steps<R, SUB>(v, c0, c1, r_mul(c2, u), c3, u, h, substeps);
                s_io[lane][kk] = (double)v;
            }
            __syncwarp();
            double *g = pred + first * W + k0;
            if (nc == R6_CH)
                for (int e = lane; e < nrows * R6_CH; e += 32) g[(int64_t)(e >> 4) * W + (e & 15)] = s_io[e >> 4][e & 15];
            else
                for (int e = lane; e < nrows * nc; e += 32) g[(int64_t)(e / nc) * W + (e % nc)] = s_io[e / nc][e % nc];
            __syncwarp();
        }
    }
}

// generic kernel (any W): thread per row; predictions staged in shared memory (odd pitch: conflict-free) and written
// back as one contiguous, fully coalesced chunk per CTA.
constexpr int RP = 128;

template <typename R>
__global__ void __launch_bounds__(RP)
ode_rollout_kernel(int64_t rows, int W, double dt, int substeps, const double *__restrict__ x0,
                   const double *__restrict__ static_feature, const uint8_t *__restrict__ codes,
                   const double *__restrict__ coefs, int per_row, double drop_below, const double *__restrict__ dts,
                   int dts_per_row, double *__restrict__ pred)
{
    const R h_uniform = (R)(dt / substeps);
    extern __shared__ double s_out[];  // [RP][W | 1]
    __shared__ double s_coef[16];
    const int pitch = W | 1;
    const int tid = threadIdx.x;
    if (!per_row && tid < 16) {
        const double c = coefs[tid];
        s_coef[tid] = (fabs(c) > drop_below) ? c : 0.0;
    }
    __syncthreads();
    const int64_t ntiles = (rows + RP - 1) / RP;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t first = tile * RP;
        const int64_t r = first + tid;
        if (r < rows) {
            R c[4][4];
            if (per_row) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const double v = coefs[r * 16 + j];
                    c[j >> 2][j & 3] = (R)((fabs(v) > drop_below) ? v : 0.0);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) c[j >> 2][j & 3] = (R)s_coef[j];
            }
            R v = (R)x0[r];
            const R u = (R)static_feature[r];
            const uint8_t *cr = codes + r * W;
            const double *dr = dts ? dts + (dts_per_row ? r * W : 0) : nullptr;
            for (int k = 0; k < W; ++k) {
                const R h = dr ? (R)(dr[k] / substeps) : h_uniform;
                v = euler_interval<R>(v, c, cr[k] & 3, u, h, substeps);
                s_out[tid * pitch + k] = (double)v;
            }
        }
        __syncthreads();
        const int nrows = (int)((rows - first < RP) ? (rows - first) : RP);
        double *dst = pred + first * W;
        for (int e = tid; e < nrows * W; e += RP) dst[e] = s_out[(e / W) * pitch + (e % W)];
        __syncthreads();
    }
}

// ---- treatment codes -------------------------------------------------------------------------------
__global__ void treatment_codes_kernel(int64_t rows, int W, int pitch, const double *__restrict__ chemo,
                                       const double *__restrict__ radio, uint8_t *__restrict__ codes)
{
    const int64_t total = rows * W;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / W;
        const int k = (int)(e - r * W);
        const double c = chemo[r * pitch + k], d = radio[r * pitch + k];
        codes[e] = (uint8_t)((c != 0.0 ? 1 : 0) + (d != 0.0 ? 2 : 0));
    }
}

// ---- masked squared errors -------------------------------------------------------------------------
// one warp per row, lanes over columns; per-lane column partials, block combine, ordered grid combine.
constexpr int MSE_MAX_W = 128;
struct MseWorkspaceHdr {
    unsigned int ticket[32];
};

__global__ void __launch_bounds__(256)
masked_se_kernel(int64_t rows, int W, const double *__restrict__ pred, const double *__restrict__ target,
                 const int *__restrict__ active_len, double *__restrict__ sums, double *__restrict__ partials,
                 unsigned int *__restrict__ ticket)
{
    __shared__ double s_acc[8][3 * MSE_MAX_W];
    __shared__ unsigned int s_is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    constexpr int SL = MSE_MAX_W / 32;  // column slots per lane
    double se[SL], cnt[SL], last[SL];
#pragma unroll
    for (int j = 0; j < SL; ++j) se[j] = cnt[j] = last[j] = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * nwarps + warp; r < rows; r += (int64_t)gridDim.x * nwarps) {
        int L = active_len[r];
        if (L > W) L = W;
#pragma unroll
        for (int j = 0; j < SL; ++j) {
            const int k = lane + 32 * j;
            if (k < L) {
                const double d = pred[r * W + k] - target[r * W + k];
                const double e = d * d;
                se[j] += e;
                cnt[j] += 1.0;
                if (k == L - 1) last[j] += e;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < SL; ++j) {
        const int k = lane + 32 * j;
        s_acc[warp][k] = se[j];
        s_acc[warp][MSE_MAX_W + k] = cnt[j];
        s_acc[warp][2 * MSE_MAX_W + k] = last[j];
    }
    __syncthreads();
    const int nvals = 3 * MSE_MAX_W;
    for (int j = tid; j < nvals; j += blockDim.x) {
        double v = 0.0;
        for (int w = 0; w < nwarps; ++w) v += s_acc[w][j];
        partials[(size_t)blockIdx.x * nvals + j] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        s_is_last = (t == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_is_last) {
        __threadfence();
        for (int j = tid; j < nvals; j += blockDim.x) {
            double v = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) v += __ldcg(&partials[(size_t)b * nvals + j]);
            const int which = j / MSE_MAX_W, k = j % MSE_MAX_W;
            if (k < W) sums[which * W + k] = v;
            s_acc[0][j] = v;
        }
        __syncthreads();
        if (tid == 0) {
            double tl = 0.0, cl = 0.0;
            for (int k = 0; k < W; ++k) tl += s_acc[0][2 * MSE_MAX_W + k];
            // number of rows with a last active entry = rows with L >= 1 = active count of column 0
            cl = s_acc[0][MSE_MAX_W + 0];
            sums[3 * W] = tl;
            sums[3 * W + 1] = cl;
            *ticket = 0u;
        }
    }
}

}  // namespace b200i

using namespace b200i;

extern "C" int b200i_stlsq_population(const double *stats, double threshold, double alpha, int32_t max_iter,
                                      double *coefs, int32_t *support, void *stream)
{
    B200I_REQUIRE(stats && coefs && support, B200I_E_ARG, "stlsq_population: NULL argument");
    B200I_REQUIRE(threshold >= 0 && alpha >= 0 && max_iter >= 1, B200I_E_ARG,
                  "stlsq_population: threshold/alpha must be >= 0 and max_iter >= 1");
    stlsq_population_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(stats, threshold, alpha, max_iter, coefs,
                                                                             support);
    return check_cuda(cudaGetLastError(), "stlsq_population launch");
}

extern "C" int b200i_stlsq_joint(const double *stats, double threshold, double alpha, int32_t max_iter, double drop_below,
                                 double *coefs11, int32_t *support11, double *coefs44, void *stream)
{
    B200I_REQUIRE(stats && coefs11 && support11 && coefs44, B200I_E_ARG, "stlsq_joint: NULL argument");
    B200I_REQUIRE(threshold >= 0 && alpha >= 0 && max_iter >= 1, B200I_E_ARG, "stlsq_joint: bad scalar argument");
    stlsq_joint_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(stats, threshold, alpha, max_iter, drop_below,
                                                                        coefs11, support11, coefs44);
    return check_cuda(cudaGetLastError(), "stlsq_joint launch");
}

static int ode_rollout_impl(bool f32, int64_t rows, int32_t W, double dt, int32_t substeps, const double *x0,
                            const double *static_feature, const uint8_t *codes, const double *coefs,
                            int32_t coefs_per_row, double drop_below, const double *dts, int32_t dts_per_row, double *pred,
                            void *stream)
{
    B200I_REQUIRE(rows >= 0 && x0 && static_feature && codes && coefs && pred, B200I_E_ARG,
                  "ode_rollout: NULL argument or negative rows");
    B200I_REQUIRE(W >= 1 && W <= 2048 && substeps >= 1, B200I_E_UNSUPPORTED, "ode_rollout: W=%d substeps=%d", W, substeps);
    B200I_REQUIRE(dts != nullptr || dt > 0, B200I_E_ARG, "ode_rollout: dt must be positive when no interval lengths are given");
    if (rows == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (W <= R6_MAXW) {
        const int code_bytes = (32 * W + 15) & ~15;
        const int smem = R6_WARPS * (32 * (R6_CH + 1) * 8 + (coefs_per_row ? 32 * 17 * 8 : 0) + code_bytes);
        const bool sub5 = substeps == 5;      // STEPS_FOR_DT of the reference (pkpd/utils.py:40): unrolled build
        const void *kern = f32 ? (sub5 ? reinterpret_cast<const void *>(ode_rollout_tiled_kernel<float, 5>)
                                       : reinterpret_cast<const void *>(ode_rollout_tiled_kernel<float, 0>))
                               : (sub5 ? reinterpret_cast<const void *>(ode_rollout_tiled_kernel<double, 5>)
                                       : reinterpret_cast<const void *>(ode_rollout_tiled_kernel<double, 0>));
        int per_sm = 1;
        {
            int rc0 = ensure_dyn_smem(kern, smem, R6_WARPS * 32, &per_sm);
            if (rc0) return rc0;
        }
        if (per_sm < 1) per_sm = 1;
        const int64_t ntiles = (rows + 31) / 32;
        int64_t grid = (ntiles + R6_WARPS - 1) / R6_WARPS;
        const int64_t cap = (int64_t)num_sms() * per_sm;
        if (grid > cap) grid = cap;
        void *args[] = {&rows, &W, &dt, &substeps, &x0, &static_feature, &codes, &coefs, &coefs_per_row, &drop_below, &dts,
                        &dts_per_row, &pred};
        B200I_CUDA(cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(R6_WARPS * 32), args, (size_t)smem, st));
        return check_cuda(cudaGetLastError(), "ode_rollout launch");
    }
    const size_t smem = (size_t)RP * (W | 1) * sizeof(double);
    B200I_REQUIRE(smem <= 200 * 1024, B200I_E_UNSUPPORTED, "ode_rollout: W=%d too wide", W);
    const void *kern = f32 ? reinterpret_cast<const void *>(ode_rollout_kernel<float>)
                           : reinterpret_cast<const void *>(ode_rollout_kernel<double>);
    int per_sm = 1;
    {
        int rc0 = ensure_dyn_smem(kern, (int)smem, RP, &per_sm);
        if (rc0) return rc0;
    }
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (rows + RP - 1) / RP;
    const int64_t cap = (int64_t)num_sms() * per_sm;
    if (grid > cap) grid = cap;
    if (f32)
        ode_rollout_kernel<float><<<(unsigned)grid, RP, smem, st>>>(rows, W, dt, substeps, x0, static_feature, codes, coefs,
                                                                     coefs_per_row, drop_below, dts, dts_per_row, pred);
    else
        ode_rollout_kernel<double><<<(unsigned)grid, RP, smem, st>>>(rows, W, dt, substeps, x0, static_feature, codes, coefs,
                                                                      coefs_per_row, drop_below, dts, dts_per_row, pred);
    return check_cuda(cudaGetLastError(), "ode_rollout launch");
}

extern "C" int b200i_ode_rollout(int64_t rows, int32_t W, double dt, int32_t substeps, const double *x0,
                                 const double *static_feature, const uint8_t *codes, const double *coefs,
                                 int32_t coefs_per_row, double drop_below, double *pred, void *stream)
{
    return ode_rollout_impl(false, rows, W, dt, substeps, x0, static_feature, codes, coefs, coefs_per_row, drop_below,
                            nullptr, 0, pred, stream);
}

extern "C" int b200i_ode_rollout_f32(int64_t rows, int32_t W, double dt, int32_t substeps, const double *x0,
                                     const double *static_feature, const uint8_t *codes, const double *coefs,
                                     int32_t coefs_per_row, double drop_below, double *pred, void *stream)
{
    return ode_rollout_impl(true, rows, W, dt, substeps, x0, static_feature, codes, coefs, coefs_per_row, drop_below,
                            nullptr, 0, pred, stream);
}

extern "C" int b200i_ode_rollout_dts(int64_t rows, int32_t W, int32_t substeps, const double *x0,
                                     const double *static_feature, const uint8_t *codes, const double *coefs,
                                     int32_t coefs_per_row, double drop_below, const double *dts, int32_t dts_per_row,
                                     int32_t fp32, double *pred, void *stream)
{
    B200I_REQUIRE(dts != nullptr, B200I_E_ARG, "ode_rollout_dts: dts is NULL");
    return ode_rollout_impl(fp32 != 0, rows, W, 0.0, substeps, x0, static_feature, codes, coefs, coefs_per_row, drop_below,
                            dts, dts_per_row, pred, stream);
}

extern "C" int b200i_treatment_codes(int64_t rows, int32_t W, int32_t row_pitch, const double *chemo_application,
                                     const double *radio_application, uint8_t *codes, void *stream)
{
    B200I_REQUIRE(rows >= 0 && W >= 1 && row_pitch >= W && chemo_application && radio_application && codes, B200I_E_ARG,
                  "treatment_codes: bad argument");
    if (rows == 0) return 0;
    int64_t grid = (rows * W + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (grid > cap) grid = cap;
    treatment_codes_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        rows, W, row_pitch, chemo_application, radio_application, codes);
    return check_cuda(cudaGetLastError(), "treatment_codes launch");
}

extern "C" int64_t b200i_masked_se_workspace_bytes(void)
{
    return (int64_t)(256 + (size_t)2048 * 3 * MSE_MAX_W * sizeof(double));
}

extern "C" int b200i_masked_se(int64_t rows, int32_t W, const double *pred, const double *target,
                               const int32_t *active_len, double *sums, void *workspace, void *stream)
{
    B200I_REQUIRE(rows >= 0 && pred && target && active_len && sums && workspace, B200I_E_ARG,
                  "masked_se: NULL argument or negative rows");
    B200I_REQUIRE(W >= 1 && W <= MSE_MAX_W, B200I_E_UNSUPPORTED, "masked_se: W=%d outside [1,%d]", W, MSE_MAX_W);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B200I_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (3 * W + 2), st));
    B200I_CUDA(cudaMemsetAsync(workspace, 0, 256, st));
    if (rows == 0) return 0;
    int64_t grid = (rows + 7) / 8;
    int64_t cap = (int64_t)num_sms() * 8;
    if (cap > 2048) cap = 2048;
    if (grid > cap) grid = cap;
    unsigned int *ticket = static_cast<unsigned int *>(workspace);
    double *partials = reinterpret_cast<double *>(static_cast<uint8_t *>(workspace) + 256);
    masked_se_kernel<<<(unsigned)grid, 256, 0, st>>>(rows, W, pred, target, active_len, sums, partials, ticket);
    return check_cuda(cudaGetLastError(), "masked_se launch");
}
