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
//   poly_tsqr_kernel      16-lane workers, one treatment each; lane j owns column j of the worker's R factor
//                         (registers) and element j of every sample row.  Rows are buffered 16 at a time in shared
//                         memory and folded into R by 16 Householder reflections of [R; B] (column-wise backward
//                         stable, so the graded columns keep their relative accuracy); the workers of a CTA and then
//                         the CTAs are merged the same way (a triangular factor is a block of 16 sample rows).
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
constexpr int TSQR_MAX_CTAS = 444;

// exponents (a, b) of x^a u^b in sklearn's PolynomialFeatures order: by total degree, then lexicographic with x first
__constant__ int8_t c_poly_a[PT] = {0, 1, 0, 2, 1, 0, 3, 2, 1, 0, 4, 3, 2, 1, 0};
__constant__ int8_t c_poly_b[PT] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4};

__device__ __forceinline__ double ipow4(double x, int e)
{
    double r = 1.0;
    for (int k = 0; k < e; ++k) r *= x;
    return r;
}

// Block update of a warp's R factor (shared memory, row-major 16 x 16).  Lane l owns column j = l & 15 of R and of one
// 16-row block of the 32 buffered sample rows: b[r] = B[16 * (l >> 4) + r][j] (registers).  [R; B] (48 x 16) is brought
// back to triangular form by 16 Householder reflections; reflection k only involves R[k][:] and B (R's other rows are
// zero in column k); its vector (v0, B[:, k]) lives in lanes k and 16 + k, which publish it in shared memory (vk);
// every lane updates its half column with fused multiply-add chains, the two halves of a dot product meet through one
// shuffle.  ~50 warp instructions per sample row and ~75 registers (three CTAs per SM); the row-by-row Givens version
// of the first cut needed ~2000 instructions per row of a 16-lane group, with a square root and two divisions per
// rotation.
constexpr int VK_HALF = PQ + 2;        // the halves' copies of the reflector sit in different banks
__device__ __forceinline__ void householder_fold(double *Rs, const double *Bs, double *vk, int lane)
{
    const int col = lane & 15, half = lane >> 4;
    double b[PQ];
#pragma unroll
    for (int r = 0; r < PQ; ++r) b[r] = Bs[(half * PQ + r) * PQ + col];
    double *myvk = vk + half * VK_HALF;
#pragma unroll
    for (int k = 0; k < PQ; ++k) {
        double sg0 = 0.0, sg1 = 0.0;
#pragma unroll
        for (int r = 0; r < PQ; r += 2) {
            sg0 = fma(b[r], b[r], sg0);
            sg1 = fma(b[r + 1], b[r + 1], sg1);
        }
        double sig_k = __shfl_sync(0xffffffffu, sg0 + sg1, k, 16);
        sig_k += __shfl_xor_sync(0xffffffffu, sig_k, 16);
        if (sig_k == 0.0) continue;                       // nothing below the diagonal in column k (warp-uniform)
        if (col == k) {
#pragma unroll
            for (int r = 0; r < PQ; ++r) myvk[r] = b[r];
        }
        const double rk = Rs[k * PQ + col], alpha = Rs[k * PQ + k];
        __syncwarp();
        const double nrm = sqrt(fma(alpha, alpha, sig_k));
        const double beta = alpha >= 0.0 ? -nrm : nrm;
        const double v0 = alpha - beta;                   // no cancellation: alpha and -beta share their sign
        const double tau = 2.0 / fma(v0, v0, sig_k);
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int r = 0; r < PQ; r += 2) {
            d0 = fma(myvk[r], b[r], d0);
            d1 = fma(myvk[r + 1], b[r + 1], d1);
        }
        double d = d0 + d1;
        d += __shfl_xor_sync(0xffffffffu, d, 16);
        const double w = tau * fma(v0, rk, d);
        if (col > k) {
            if (half == 0) Rs[k * PQ + col] = fma(-w, v0, rk);
#pragma unroll
            for (int r = 0; r < PQ; ++r) b[r] = fma(-w, myvk[r], b[r]);
        } else if (col == k) {
            if (half == 0) Rs[k * PQ + col] = beta;
#pragma unroll
            for (int r = 0; r < PQ; ++r) b[r] = 0.0;
        }
        __syncwarp();                                     // vk and R[k][k] are free again
    }
}

constexpr int TSQR_WARPS = 8;
// Warp -> (treatment, patient slot).  The untreated steps carry ~45 % of the sample rows of a cancer_sim cohort, chemo
// only and radio only ~22 % each, both ~11 %: three, two, two and one warp keep the warps' row counts within 1.4x.
__constant__ int8_t c_tsqr_a[TSQR_WARPS] = {0, 1, 2, 3, 0, 1, 2, 0};
__constant__ int8_t c_tsqr_slot[TSQR_WARPS] = {0, 0, 0, 0, 1, 1, 1, 2};
__constant__ int8_t c_tsqr_nslots[4] = {3, 2, 2, 1};
constexpr int TSQR_ROWS = 32;          // buffered sample rows per fold
constexpr int TSQR_WARP_DOUBLES = PQ * PQ + TSQR_ROWS * PQ + 2 * VK_HALF;   // R, B, vk

// A warp serves one treatment on its share of the patients: it stages a patient's volumes
// and treatment codes in shared memory, finds the steps of its treatment 32 at a time (ballot), buffers their sample
// rows (32) and folds full buffers into its R factor.
__global__ void __launch_bounds__(TSQR_WARPS * 32, 3)
poly_tsqr_kernel(int64_t n, int T, double fd_dt, const double *__restrict__ vol, const double *__restrict__ chemo,
                 const double *__restrict__ radio, const double *__restrict__ seq, const double *__restrict__ stat,
                 double *__restrict__ scratch, unsigned long long *__restrict__ counts)
{
    extern __shared__ __align__(16) double s_tsqr[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, col = lane & 15, half = lane >> 4;
    const int a = c_tsqr_a[warp], nslots = c_tsqr_nslots[a];
    double *Rs = s_tsqr + (size_t)warp * TSQR_WARP_DOUBLES;                              // [16][16]
    double *Bs = Rs + PQ * PQ;                                                           // [32][16]
    double *vk = Bs + TSQR_ROWS * PQ;
    double *xs = s_tsqr + (size_t)TSQR_WARPS * TSQR_WARP_DOUBLES + (size_t)warp * (T + 1);   // volumes of the patient
    uint8_t *cs = reinterpret_cast<uint8_t *>(s_tsqr + (size_t)TSQR_WARPS * (TSQR_WARP_DOUBLES + T + 1)) + (size_t)warp * (T + 1);
    for (int e = lane; e < PQ * PQ; e += 32) Rs[e] = 0.0;
    const int ea = col < PT ? c_poly_a[col] : 0, eb = col < PT ? c_poly_b[col] : 0;
    const bool unit_dt = fd_dt == 1.0;
    unsigned long long rows = 0ull;
    int cnt = 0;
    const int64_t stride = (int64_t)gridDim.x * nslots;
    for (int64_t p = (int64_t)blockIdx.x * nslots + c_tsqr_slot[warp]; p < n; p += stride) {
        const double *x = vol + p * T, *ch = chemo + p * T, *ra = radio + p * T;
        const int L = min((int)seq[p], T - 1);
        if (L <= 0) continue;                             // no step, no sample (warp-uniform)
        const double ub = ipow4(stat[p], eb);
        __syncwarp();
        for (int k = lane; k <= L; k += 32) xs[k] = x[k];
        for (int k = lane; k < L; k += 32) cs[k] = (uint8_t)((ch[k] != 0.0 ? 1 : 0) + (ra[k] != 0.0 ? 2 : 0));
        if (lane == 0) cs[L] = 0xff;                      // the last step always closes its snippet
        __syncwarp();
        for (int base = 0; base < L; base += 32) {
            const int kk = base + lane;
            unsigned m = __ballot_sync(0xffffffffu, kk < L && cs[kk] == a);
            while (m) {
                const int i = base + __ffs(m) - 1;
                m &= m - 1;
                const double x0 = xs[i], x1 = xs[i + 1];
                const double xd = unit_dt ? (x1 - x0) : (x1 - x0) / fd_dt;
                const int nrow = (cs[i + 1] != cs[i]) ? 2 : 1;   // the snippet's end point carries the backward difference
                for (int e = 0; e < nrow; ++e) {
                    // this lane's monomial x^ea u^eb: the powers by repeated multiplication (1 * x * x ...), chosen by
                    // selects -- a loop over the lane's exponent diverges inside the warp and cost 70 instructions a row
                    const double x = e ? x1 : x0;
                    const double p2 = x * x, p3 = p2 * x, p4 = p3 * x;
                    const double pw = ea == 0 ? 1.0 : (ea == 1 ? x : (ea == 2 ? p2 : (ea == 3 ? p3 : p4)));
                    if ((cnt >> 4) == half) Bs[cnt * PQ + col] = col < PT ? pw * ub : xd;
                    if (++cnt == TSQR_ROWS) {
                        __syncwarp();
                        householder_fold(Rs, Bs, vk, lane);
                        cnt = 0;
                    }
                }
                rows += nrow;
            }
        }
    }
    if (cnt) {
        for (int r = cnt; r < TSQR_ROWS; ++r)
            if ((r >> 4) == half) Bs[r * PQ + col] = 0.0;
        __syncwarp();
        householder_fold(Rs, Bs, vk, lane);
    }
    if (lane == 0 && rows) atomicAdd(&counts[a], rows);
    // merge the warps of a treatment: a parked factor is a block of 16 sample rows (rows 16..31 of the block zero)
    __syncwarp();
    if (warp >= 4)
        for (int e = lane; e < PQ * PQ; e += 32) { Bs[e] = Rs[e]; Bs[PQ * PQ + e] = 0.0; }
    __syncthreads();
    if (warp < 4) {       // warps 0..3 lead treatments 0..3
        for (int o = 4; o < TSQR_WARPS; ++o)
            if (c_tsqr_a[o] == a) householder_fold(Rs, s_tsqr + (size_t)o * TSQR_WARP_DOUBLES + PQ * PQ, vk, lane);
        double *g = scratch + ((size_t)blockIdx.x * 4 + warp) * PQ * PQ;
        for (int e = lane; e < PQ * PQ; e += 32) g[e] = Rs[e];
    }
}

// one CTA, warp a = treatment a: folds the CTA factors two at a time (rows 0..15 and 16..31 of the block)
__global__ void __launch_bounds__(128)
poly_tsqr_merge_kernel(int nctas, const double *__restrict__ scratch, const unsigned long long *__restrict__ counts,
                       double *__restrict__ r_out)
{
    __shared__ double s_w[4][TSQR_WARP_DOUBLES];
    const int tid = threadIdx.x, a = tid >> 5, lane = tid & 31;
    double *Rs = s_w[a], *Bs = Rs + PQ * PQ, *vk = Bs + TSQR_ROWS * PQ;
    for (int e = lane; e < PQ * PQ; e += 32) Rs[e] = scratch[(size_t)a * PQ * PQ + e];
    for (int c = 1; c < nctas; c += 2) {
        __syncwarp();
        for (int e = lane; e < 2 * PQ * PQ; e += 32) {
            const int cc = c + e / (PQ * PQ);
            Bs[e] = (cc < nctas) ? scratch[((size_t)cc * 4 + a) * PQ * PQ + e % (PQ * PQ)] : 0.0;
        }
        __syncwarp();
        householder_fold(Rs, Bs, vk, lane);
    }
    __syncwarp();
    for (int e = lane; e < PQ * PQ; e += 32) r_out[(size_t)a * PQ * PQ + e] = Rs[e];
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
    const double count = r_factors[4 * PQ * PQ + a];
    const bool empty = count == 0.0;
    if (!(rcond > 0.0)) rcond = fmax(count, (double)PT) * 2.220446049250313e-16;   // numpy.linalg.lstsq's eps * max(M, N)
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
    B200I_REQUIRE(n >= 0 && workspace && r_out, B200I_E_ARG, "poly_tsqr: NULL workspace / output or negative n");
    B200I_REQUIRE(n == 0 || (cancer_volume && chemo_application && radio_application && sequence_lengths && static_feature),
                  B200I_E_ARG, "poly_tsqr: NULL argument");   // the arrays of an empty cohort may be NULL
    B200I_REQUIRE(T >= 2, B200I_E_UNSUPPORTED, "poly_tsqr: T=%d < 2", T);
    B200I_REQUIRE(fd_dt > 0.0, B200I_E_ARG, "poly_tsqr: fd_dt must be positive");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long *counts = static_cast<unsigned long long *>(workspace);
    double *scratch = reinterpret_cast<double *>(static_cast<uint8_t *>(workspace) + 64);
    B200I_CUDA(cudaMemsetAsync(counts, 0, 64, st));
    B200I_REQUIRE(T <= 1024, B200I_E_UNSUPPORTED, "poly_tsqr: T=%d > 1024", T);
    // a CTA takes ~128 patients before another CTA pays for its merge
    int64_t grid = (n + 127) / 128;
    if (grid < 1) grid = 1;
    const int64_t cap = (3 * (int64_t)num_sms() < TSQR_MAX_CTAS) ? 3 * (int64_t)num_sms() : TSQR_MAX_CTAS;
    if (grid > cap) grid = cap;
    const int smem = TSQR_WARPS * ((TSQR_WARP_DOUBLES + T + 1) * (int)sizeof(double) + ((T + 1 + 7) & ~7));
    {
        int per_sm = 1;
        int rc0 = ensure_dyn_smem(reinterpret_cast<const void *>(poly_tsqr_kernel), smem, TSQR_WARPS * 32, &per_sm);
        if (rc0) return rc0;
    }
    poly_tsqr_kernel<<<(unsigned)grid, TSQR_WARPS * 32, smem, st>>>(n, T, fd_dt, cancer_volume, chemo_application,
                                                                    radio_application, sequence_lengths, static_feature,
                                                                    scratch, counts);
    B200I_CUDA(cudaGetLastError());
    poly_tsqr_merge_kernel<<<1, 128, 0, st>>>((int)grid, scratch, counts, r_out);
    return check_cuda(cudaGetLastError(), "poly_tsqr launch");
}

extern "C" int b200i_poly_stlsq(const double *r_factors, double threshold, double alpha, int32_t max_iter, double rcond,
                                double *coefs_out, int32_t *support_out, void *stream)
{
    B200I_REQUIRE(r_factors && coefs_out && support_out, B200I_E_ARG, "poly_stlsq: NULL argument");
    B200I_REQUIRE(alpha >= 0.0 && max_iter >= 1, B200I_E_ARG, "poly_stlsq: alpha must be >= 0 and max_iter >= 1");
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
