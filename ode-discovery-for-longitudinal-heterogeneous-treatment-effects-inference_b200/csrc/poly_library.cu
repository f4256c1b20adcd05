// poly_library.cu -- the "more complex basis functions" ablation of the population SINDy fit
// (model.ablation_more_complex_basis_functions, libs_m/ct/src/models/sindy.py:185-186):
// PolynomialLibrary(degree=4, interaction_only=False) over [x0 = volume, u0 = patient type], 15 monomials per treatment.
//
// Why not the Gram path of K4/K5: with x up to ~1150 the columns span 1 .. x^4 (sigma_max ~ 1e13) and u takes three
// values, so u^0..u^4 and x u^0..x u^3 are rank deficient by construction (rank 12 of 15 on every cancer_sim cohort).
// pysindy's answer is nevertheless well defined -- ridge steps pick the support, the final un-biasing is scipy's
// minimum-norm least squares (SVD) on the design matrix -- but it cannot be recovered from Theta^T Theta in FP64:
// the squared condition number buries the two smallest genuine singular values.  So the statistics here are the
// R factor of [Theta | xdot] per treatment (16 x 16), built by a tall-skinny QR:
//
//   poly_tsqr_kernel      16-lane workers; lane j owns column j of the worker's four R factors (shared memory) and
//                         element j of the incoming sample row; every sample row is rotated in with 16 Givens
//                         rotations (column-wise backward stable, so graded columns keep their relative accuracy);
//                         the workers of a CTA are merged the same way (rows of a triangular factor are sample rows
//                         whose leading zeros are skipped), one CTA factor per treatment goes to scratch.
//   poly_tsqr_merge_kernel one CTA, one 16-lane group per treatment: merges the CTA factors in a fixed order.
//   poly_stlsq_kernel     one warp per treatment: pysindy STLSQ on the R factor.  Ridge and minimum-norm solutions come
//                         from a one-sided Jacobi SVD of R[:, support] (lane r owns row r):
//                         ridge  c = V diag(1 / (s^2 + alpha)) (A V)^T z      == (Theta^T Theta + alpha I)^-1 Theta^T xdot
//                         unbias c = V diag(s > rcond s_max ? 1 / s^2 : 0) (A V)^T z   == scipy.linalg.lstsq (gelsd)
//   poly_rollout_kernel   K6 for the polynomial ODE: per row and treatment the quartic in x with coefficients
//                         q_k(u) = sum_b c[k,b] u^b; explicit Euler, `substeps` per interval; warp tile staging of the
//                         codes and the predictions as in ode_rollout_tiled_kernel.
//
// Trajectories, finite differences and sample rows exactly as theta_gram cuts them (per-treatment snippets; every
// snippet contributes its samples plus its end point with the backward difference).
#include "common.cuh"

namespace b200i {

constexpr int PQ = 16;                 // 15 monomials + the derivative column
constexpr int PT = B200I_POLY_TERMS;   // 15
constexpr int TSQR_WORKERS = 8;        // 16-lane workers per CTA (256 threads would be 16; 128 threads = 8)
constexpr int TSQR_MAX_CTAS = 296;

// exponents (a, b) of x^a u^b in sklearn's PolynomialFeatures order: by total degree, then lexicographic with x first
__constant__ int8_t c_poly_a[PT] = {0, 1, 0, 2, 1, 0, 3, 2, 1, 0, 4, 3, 2, 1, 0};
__constant__ int8_t c_poly_b[PT] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4};

__device__ __forceinline__ double ipow4(double x, int e)
{
    double r = 1.0;
    for (int k = 0; k < e; ++k) r *= x;
    return r;
}

// Rotate the row held across the 16 lanes of a group (lane j: v_j) into the upper-triangular R (row-major 16x16 in
// shared memory, lane j touches column j only).  `first`: leading zeros of the row (rotations before it are skipped).
__device__ __forceinline__ void givens_insert(double *R, double v, int lane16, int first, unsigned gmask)
{
    for (int k = first; k < PQ; ++k) {
        const double t = R[k * PQ + lane16];
        const double rkk = __shfl_sync(gmask, t, k, 16);
        const double vk = __shfl_sync(gmask, v, k, 16);
        if (vk != 0.0) {                                  // uniform within the group
            const double r = sqrt(rkk * rkk + vk * vk);
            const double c = rkk / r, s = vk / r;
            if (lane16 >= k) {
                R[k * PQ + lane16] = c * t + s * v;
                v = c * v - s * t;
            }
            if (lane16 == k) v = 0.0;
        }
    }
}

__global__ void __launch_bounds__(TSQR_WORKERS * 16)
poly_tsqr_kernel(int64_t n, int T, double fd_dt, const double *__restrict__ vol, const double *__restrict__ chemo,
                 const double *__restrict__ radio, const double *__restrict__ seq, const double *__restrict__ stat,
                 double *__restrict__ scratch, unsigned long long *__restrict__ counts)
{
    extern __shared__ __align__(16) double s_tsqr[];
    double(*s_R)[4][PQ * PQ] = reinterpret_cast<double(*)[4][PQ * PQ]>(s_tsqr);
    __shared__ unsigned long long s_cnt[4];
    const int tid = threadIdx.x, worker = tid >> 4, lane16 = tid & 15;
    const unsigned gmask = 0xffffu << (16 * ((tid >> 4) & 1));
    for (int e = tid; e < TSQR_WORKERS * 4 * PQ * PQ; e += blockDim.x) s_tsqr[e] = 0.0;
    if (tid < 4) s_cnt[tid] = 0ull;
    __syncthreads();
    const int ea = lane16 < PT ? c_poly_a[lane16] : 0, eb = lane16 < PT ? c_poly_b[lane16] : 0;
    unsigned long long my_cnt[4] = {0ull, 0ull, 0ull, 0ull};
    const int64_t stride = (int64_t)gridDim.x * TSQR_WORKERS;
    for (int64_t p = (int64_t)blockIdx.x * TSQR_WORKERS + worker; p < n; p += stride) {
        const double *x = vol + p * T, *ch = chemo + p * T, *ra = radio + p * T;
        const int L = min((int)seq[p], T - 1);
        const double u = stat[p];
        const double ub = ipow4(u, eb);
        for (int i = 0; i < L; ++i) {
            const int code = (ch[i] != 0.0 ? 1 : 0) + (ra[i] != 0.0 ? 2 : 0);
            const double x0 = x[i], x1 = x[i + 1];
            const double xd = (x1 - x0) / fd_dt;
            double *R = &s_R[worker][code][0];
            givens_insert(R, lane16 < PT ? ipow4(x0, ea) * ub : xd, lane16, 0, gmask);
            int rows = 1;
            const bool last = (i == L - 1) || (ch[i + 1] != ch[i]) || (ra[i + 1] != ra[i]);
            if (last) {   // the snippet's end point with the backward difference
                givens_insert(R, lane16 < PT ? ipow4(x1, ea) * ub : xd, lane16, 0, gmask);
                rows = 2;
            }
            if (lane16 == 0) my_cnt[code] += rows;
        }
    }
    if (lane16 == 0)
        for (int a = 0; a < 4; ++a)
            if (my_cnt[a]) atomicAdd(&s_cnt[a], my_cnt[a]);
    __syncthreads();
    // merge the workers' factors: group a (a < 4) owns treatment a and folds the other workers' rows into its own
    if (worker < 4) {
        double *R = &s_R[worker][worker][0];
        for (int o = 0; o < TSQR_WORKERS; ++o) {
            if (o == worker) continue;
            const double *S = &s_R[o][worker][0];
            for (int k = 0; k < PQ; ++k) givens_insert(R, lane16 >= k ? S[k * PQ + lane16] : 0.0, lane16, k, gmask);
        }
        double *g = scratch + ((size_t)blockIdx.x * 4 + worker) * PQ * PQ;
        for (int k = 0; k < PQ; ++k) g[k * PQ + lane16] = R[k * PQ + lane16];
    }
    if (tid < 4 && s_cnt[tid]) atomicAdd(&counts[tid], s_cnt[tid]);
}

__global__ void __launch_bounds__(64)
poly_tsqr_merge_kernel(int nctas, const double *__restrict__ scratch, const unsigned long long *__restrict__ counts,
                       double *__restrict__ r_out)
{
    __shared__ double s_R[4][PQ * PQ];
    const int tid = threadIdx.x, a = tid >> 4, lane16 = tid & 15;
    const unsigned gmask = 0xffffu << (16 * (a & 1));
    double *R = &s_R[a][0];
    for (int k = 0; k < PQ; ++k) R[k * PQ + lane16] = scratch[(size_t)a * PQ * PQ + k * PQ + lane16];
    __syncwarp(gmask);
    for (int c = 1; c < nctas; ++c) {
        const double *S = scratch + ((size_t)c * 4 + a) * PQ * PQ;
        for (int k = 0; k < PQ; ++k) givens_insert(R, lane16 >= k ? S[k * PQ + lane16] : 0.0, lane16, k, gmask);
    }
    for (int k = 0; k < PQ; ++k) r_out[(size_t)a * PQ * PQ + k * PQ + lane16] = R[k * PQ + lane16];
    if (tid < 4) r_out[4 * PQ * PQ + tid] = (double)counts[tid];
}

// ---- STLSQ on the R factor -----------------------------------------------------------------------------------
// One warp per treatment.  A = R[0:15, support] (15 x m, lane r = row r), z = R[0:15, 15].
struct PolySvd {
    double A[PT][PQ];   // working copy, columns rotated to orthogonality: A V
    double V[PT][PQ];   // accumulated right singular vectors (rows = position in the support)
};

__device__ void jacobi_svd(PolySvd &w, int m, int lane)
{
    const bool row = lane < PT;
    for (int sweep = 0; sweep < 60; ++sweep) {
        int rotated = 0;
        for (int p = 0; p < m - 1; ++p)
            for (int q = p + 1; q < m; ++q) {
                const double ap = row ? w.A[lane][p] : 0.0, aq = row ? w.A[lane][q] : 0.0;
                const double alpha = warp_sum(ap * ap), beta = warp_sum(aq * aq), gamma = warp_sum(ap * aq);
                if (alpha == 0.0 || beta == 0.0) continue;
                if (fabs(gamma) <= 1e-15 * sqrt(alpha) * sqrt(beta)) continue;
                ++rotated;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                if (row) {
                    w.A[lane][p] = c * ap - s * aq;
                    w.A[lane][q] = s * ap + c * aq;
                }
                if (lane < m) {
                    const double vp = w.V[lane][p], vq = w.V[lane][q];
                    w.V[lane][p] = c * vp - s * vq;
                    w.V[lane][q] = s * vp + c * vq;
                }
                __syncwarp();
            }
        if (!rotated) break;
    }
}

// c[support] = V diag(f(s)) (A V)^T z with f = 1 / (s^2 + alpha) (ridge >= 0) or the truncated inverse (ridge < 0).
__device__ void svd_solve(const double *R, unsigned support, double ridge, double rcond, PolySvd &w, double *coef, int lane)
{
    int idx[PT], m = 0;
    for (int j = 0; j < PT; ++j)
        if ((support >> j) & 1u) idx[m++] = j;
    if (lane < PT)
        for (int j = 0; j < m; ++j) w.A[lane][j] = (lane <= idx[j]) ? R[lane * PQ + idx[j]] : 0.0;
    if (lane < m)
        for (int j = 0; j < m; ++j) w.V[lane][j] = (lane == j) ? 1.0 : 0.0;
    __syncwarp();
    jacobi_svd(w, m, lane);
    const double z = lane < PT ? R[lane * PQ + PT] : 0.0;
    double s2[PT], wz[PT], smax2 = 0.0;
    for (int j = 0; j < m; ++j) {
        const double a = lane < PT ? w.A[lane][j] : 0.0;
        s2[j] = warp_sum(a * a);
        wz[j] = warp_sum(a * z);
        smax2 = fmax(smax2, s2[j]);
    }
    double c = 0.0;
    for (int j = 0; j < m; ++j) {
        double d;
        if (ridge >= 0.0)
            d = wz[j] / (s2[j] + ridge);
        else
            d = (sqrt(s2[j]) > rcond * sqrt(smax2)) ? wz[j] / s2[j] : 0.0;
        if (lane < m) c += w.V[lane][j] * d;
    }
    if (lane < PT) coef[lane] = 0.0;
    __syncwarp();
    if (lane < m) coef[idx[lane]] = c;
    __syncwarp();
}

__global__ void __launch_bounds__(128)
poly_stlsq_kernel(const double *__restrict__ r_factors, double threshold, double alpha, int max_iter, double rcond,
                  double *__restrict__ coefs_out, int *__restrict__ support_out)
{
    __shared__ PolySvd s_w[4];
    __shared__ double s_R[4][PQ * PQ];
    __shared__ double s_coef[4][PQ];
    const int a = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = lane; e < PQ * PQ; e += 32) s_R[a][e] = r_factors[(size_t)a * PQ * PQ + e];
    __syncwarp();
    const double *R = s_R[a];
    double *coef = s_coef[a];
    const bool empty = r_factors[4 * PQ * PQ + a] == 0.0;
    // pkpd/utils.py:274-310 (pysindy STLSQ._reduce): ridge on the support, hard threshold, stop when nothing was dropped
    // in this pass or the support repeats
    unsigned ind = (1u << PT) - 1u, last = ind;
    int n_sel = PT;
    if (lane < PQ) coef[lane] = 0.0;
    __syncwarp();
    if (!empty) {
        for (int it = 0; it < max_iter; ++it) {
            if (ind == 0u) {
                if (lane < PT) coef[lane] = 0.0;
                break;
            }
            svd_solve(R, ind, alpha, rcond, s_w[a], coef, lane);
            const bool big = lane < PT && fabs(coef[lane]) >= threshold;
            const unsigned nb = __ballot_sync(0xffffffffu, big) & ((1u << PT) - 1u);
            if (lane < PT && !big) coef[lane] = 0.0;
            __syncwarp();
            const bool same = nb == last && it > 0;
            ind = nb;
            if (__popc(ind) == n_sel || same) break;
            n_sel = __popc(ind);
            last = nb;
        }
        if (ind) svd_solve(R, ind, -1.0, rcond, s_w[a], coef, lane);   // un-bias: minimum-norm OLS on the support
    } else {
        ind = 0u;
    }
    if (lane < PT) {
        coefs_out[a * PT + lane] = coef[lane];
        support_out[a * PT + lane] = (int)((ind >> lane) & 1u);
    }
}

// ---- rollout of the polynomial ODE -----------------------------------------------------------------------------
constexpr int PR_WARPS = 4;
constexpr int PR_CH = 16;
constexpr int PR_MAXW = 128;

__global__ void __launch_bounds__(PR_WARPS * 32)
poly_rollout_kernel(int64_t rows, int W, double dt, int substeps, const double *__restrict__ x0,
                    const double *__restrict__ static_feature, const uint8_t *__restrict__ codes,
                    const double *__restrict__ coefs, double drop_below, double *__restrict__ pred)
{
    extern __shared__ __align__(16) uint8_t smem_pr[];
    __shared__ double s_c[4][PT];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int code_bytes = (32 * W + 15) & ~15;
    double(*s_io)[PR_CH + 1] = reinterpret_cast<double(*)[PR_CH + 1]>(smem_pr) + (size_t)warp * 32;
    double(*s_q)[21] = reinterpret_cast<double(*)[21]>(smem_pr + (size_t)PR_WARPS * 32 * (PR_CH + 1) * 8) + (size_t)warp * 32;
    uint8_t *s_code = smem_pr + (size_t)PR_WARPS * 32 * ((PR_CH + 1) + 21) * 8 + (size_t)warp * code_bytes;
    if (tid < 4 * PT) {
        const double c = coefs[tid];
        (&s_c[0][0])[tid] = (fabs(c) > drop_below) ? c : 0.0;
    }
    __syncthreads();
    const double h = dt / substeps;
    const int64_t ntiles = (rows + 31) / 32;
    for (int64_t tile = (int64_t)blockIdx.x * PR_WARPS + warp; tile < ntiles; tile += (int64_t)gridDim.x * PR_WARPS) {
        const int64_t first = tile * 32;
        const int nrows = (int)((rows - first < 32) ? (rows - first) : 32);
        const bool live = lane < nrows;
        __syncwarp();
        for (int e = lane; e < nrows * W; e += 32) s_code[e] = codes[first * W + e];
        double v = 0.0, u = 0.0;
        if (live) {
            v = x0[first + lane];
            u = static_feature[first + lane];
        }
        // q[a][k] = sum_b c[a][(k, b)] u^b: the row's quartic in x per treatment
        {
            double up[5] = {1.0, u, u * u, u * u * u, u * u * u * u};
            for (int a = 0; a < 4; ++a) {
                double q[5] = {0, 0, 0, 0, 0};
                for (int j = 0; j < PT; ++j) q[c_poly_a[j]] += s_c[a][j] * up[c_poly_b[j]];
                for (int k = 0; k < 5; ++k) s_q[lane][a * 5 + k] = q[k];
            }
        }
        __syncwarp();
        const uint8_t *crow = s_code + (live ? lane : 0) * W;
        for (int k0 = 0; k0 < W; k0 += PR_CH) {
            const int nc = (W - k0 < PR_CH) ? (W - k0) : PR_CH;
            for (int kk = 0; kk < nc; ++kk) {
                const double *q = &s_q[lane][5 * (crow[k0 + kk] & 3)];
                const double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4];
                for (int s = 0; s < substeps; ++s) {
                    const double f = q0 + v * (q1 + v * (q2 + v * (q3 + v * q4)));
                    v = v + h * f;
                }
                s_io[lane][kk] = v;
            }
            __syncwarp();
            double *g = pred + first * W + k0;
            for (int e = lane; e < nrows * nc; e += 32) g[(int64_t)(e / nc) * W + (e % nc)] = s_io[e / nc][e % nc];
            __syncwarp();
        }
    }
}

}  // namespace b200i

using namespace b200i;

extern "C" int64_t b200i_poly_workspace_bytes(void)
{
    return (int64_t)((size_t)TSQR_MAX_CTAS * 4 * PQ * PQ * sizeof(double) + 64);
}

extern "C" int b200i_poly_tsqr(int64_t n, int32_t T, double fd_dt, const double *cancer_volume,
                               const double *chemo_application, const double *radio_application,
                               const double *sequence_lengths, const double *static_feature, void *workspace,
                               double *r_out, void *stream)
{
    B200I_REQUIRE(n >= 0 && cancer_volume && chemo_application && radio_application && sequence_lengths && static_feature &&
                      workspace && r_out,
                  B200I_E_ARG, "poly_tsqr: NULL argument or negative n");
    B200I_REQUIRE(T >= 2, B200I_E_UNSUPPORTED, "poly_tsqr: T=%d < 2", T);
    B200I_REQUIRE(fd_dt > 0.0, B200I_E_ARG, "poly_tsqr: fd_dt must be positive");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long *counts = static_cast<unsigned long long *>(workspace);
    double *scratch = reinterpret_cast<double *>(static_cast<uint8_t *>(workspace) + 64);
    B200I_CUDA(cudaMemsetAsync(counts, 0, 64, st));
    // a CTA's 8 workers take ~32 patients each before another CTA pays for its merge
    int64_t grid = (n + TSQR_WORKERS * 32 - 1) / (TSQR_WORKERS * 32);
    if (grid < 1) grid = 1;
    const int64_t cap = (2 * (int64_t)num_sms() < TSQR_MAX_CTAS) ? 2 * (int64_t)num_sms() : TSQR_MAX_CTAS;
    if (grid > cap) grid = cap;
    const int smem = TSQR_WORKERS * 4 * PQ * PQ * (int)sizeof(double);
    {
        int per_sm = 1;
        int rc0 = ensure_dyn_smem(reinterpret_cast<const void *>(poly_tsqr_kernel), smem, TSQR_WORKERS * 16, &per_sm);
        if (rc0) return rc0;
    }
    poly_tsqr_kernel<<<(unsigned)grid, TSQR_WORKERS * 16, smem, st>>>(n, T, fd_dt, cancer_volume, chemo_application,
                                                                   radio_application, sequence_lengths, static_feature,
                                                                   scratch, counts);
    B200I_CUDA(cudaGetLastError());
    poly_tsqr_merge_kernel<<<1, 64, 0, st>>>((int)grid, scratch, counts, r_out);
    return check_cuda(cudaGetLastError(), "poly_tsqr launch");
}

extern "C" int b200i_poly_stlsq(const double *r_factors, double threshold, double alpha, int32_t max_iter, double rcond,
                                double *coefs_out, int32_t *support_out, void *stream)
{
    B200I_REQUIRE(r_factors && coefs_out && support_out, B200I_E_ARG, "poly_stlsq: NULL argument");
    B200I_REQUIRE(alpha >= 0.0 && max_iter >= 1, B200I_E_ARG, "poly_stlsq: alpha must be >= 0 and max_iter >= 1");
    if (!(rcond > 0.0)) rcond = 2.220446049250313e-16;   // scipy.linalg.lstsq(cond=None): machine epsilon
    poly_stlsq_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(r_factors, threshold, alpha, max_iter, rcond,
                                                                        coefs_out, support_out);
    return check_cuda(cudaGetLastError(), "poly_stlsq launch");
}

extern "C" int b200i_poly_rollout(int64_t rows, int32_t W, double dt, int32_t substeps, const double *x0,
                                  const double *static_feature, const uint8_t *codes, const double *coefs,
                                  double drop_below, double *pred, void *stream)
{
    B200I_REQUIRE(rows >= 0 && x0 && static_feature && codes && coefs && pred, B200I_E_ARG,
                  "poly_rollout: NULL argument or negative rows");
    B200I_REQUIRE(W >= 1 && W <= PR_MAXW, B200I_E_UNSUPPORTED, "poly_rollout: W=%d outside [1,%d]", W, PR_MAXW);
    B200I_REQUIRE(substeps >= 1 && dt > 0.0, B200I_E_ARG, "poly_rollout: substeps >= 1 and dt > 0 required");
    if (rows == 0) return 0;
    const int code_bytes = (32 * W + 15) & ~15;
    const int smem = PR_WARPS * (32 * ((PR_CH + 1) + 21) * 8 + code_bytes);
    int per_sm = 1;
    {
        int rc0 = ensure_dyn_smem(reinterpret_cast<const void *>(poly_rollout_kernel), smem, PR_WARPS * 32, &per_sm);
        if (rc0) return rc0;
    }
    if (per_sm < 1) per_sm = 1;
    const int64_t ntiles = (rows + 31) / 32;
    int64_t grid = (ntiles + PR_WARPS - 1) / PR_WARPS;
    const int64_t cap = (int64_t)num_sms() * per_sm;
    if (grid > cap) grid = cap;
    poly_rollout_kernel<<<(unsigned)grid, PR_WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(
        rows, W, dt, substeps, x0, static_feature, codes, coefs, drop_below, pred);
    return check_cuda(cudaGetLastError(), "poly_rollout launch");
}
