// sim_math.cuh -- per-patient arithmetic of the cancer PK-PD simulator, shared by K1/K2/K3.
//
// Every floating-point operation that the reference performs in numpy float64
// (cancer_simulation.py:300-349, :471-502, :671-743) is written with the explicit round-to-nearest
// intrinsics (__dadd_rn, __dmul_rn, __ddiv_rn) in the reference's evaluation order, so nvcc cannot
// contract them into FMAs and the only deviations from numpy are the <=1-2 ulp differences of
// log / exp / cbrt-vs-pow(.,1/3).
#pragma once
#include "common.cuh"
#include "fastmath.cuh"

namespace b200i {

struct Patient {
    double v0, alpha, rho, beta, beta_c, K;
    double chemo_int, radio_int, chemo_beta, radio_beta;
    bool same_sigmoid;  // chemo and radio sigmoid share (beta, intercept): one exp serves both
};

__device__ __forceinline__ Patient load_patient(const double *__restrict__ params, int64_t n, int64_t i)
{
    Patient p;
    p.v0 = __ldg(params + 0 * n + i);
    p.alpha = __ldg(params + 1 * n + i);
    p.rho = __ldg(params + 2 * n + i);
    p.beta = __ldg(params + 3 * n + i);
    p.beta_c = __ldg(params + 4 * n + i);
    p.K = __ldg(params + 5 * n + i);
    p.chemo_int = __ldg(params + 6 * n + i);
    p.radio_int = __ldg(params + 7 * n + i);
    p.chemo_beta = __ldg(params + 8 * n + i);
    p.radio_beta = __ldg(params + 9 * n + i);
    p.same_sigmoid = (p.chemo_int == p.radio_int) && (p.chemo_beta == p.radio_beta);
    return p;
}

// calc_diameter, cancer_simulation.py:38-39: ((v / (4/3*pi)) ** (1/3)) * 2.  numpy's float power
// of a negative base is NaN; cbrt would return the real root, so the sign is handled explicitly.
__device__ __forceinline__ double calc_diameter(double v, double sphere_coef)
{
    const double q = __ddiv_rn(v, sphere_coef);
    const double r = (q < 0.0) ? __longlong_as_double(0x7ff8000000000000LL) : cbrt(q);
    return __dmul_rn(r, 2.0);
}

// np.mean of n (1..MAXN<=16) register values a[0..n-1]: identity + pairwise sum exactly as numpy's
// DOUBLE_pairwise_sum (n < 8: sequential; 8 <= n <= 128: eight accumulators, tree, then tail).
template <int MAXN>
__device__ __forceinline__ double np_mean(const double (&a)[MAXN], int n)
{
    double res;
    if (MAXN == 16 && n == 16) {
        // the steady state of the simulators (window_size 15 -> 16 entries): no copies, and / 16 is an exact scaling
        res = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(a[0], a[MAXN >= 16 ? 8 : 0]), __dadd_rn(a[1], a[MAXN >= 16 ? 9 : 0])),
                                  __dadd_rn(__dadd_rn(a[2], a[MAXN >= 16 ? 10 : 0]), __dadd_rn(a[3], a[MAXN >= 16 ? 11 : 0]))),
                        __dadd_rn(__dadd_rn(__dadd_rn(a[4], a[MAXN >= 16 ? 12 : 0]), __dadd_rn(a[5], a[MAXN >= 16 ? 13 : 0])),
                                  __dadd_rn(__dadd_rn(a[6], a[MAXN >= 16 ? 14 : 0]), __dadd_rn(a[7], a[MAXN >= 16 ? 15 : 0]))));
        return __dmul_rn(__dadd_rn(0.0, res), 0.0625);
    }
    if (n < 8) {
        res = -0.0;
#pragma unroll
        for (int j = 0; j < 7; ++j)
            if (j < MAXN && j < n) res = __dadd_rn(res, a[j]);
    } else {
        double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
        if (MAXN >= 16 && n >= 16) {
            r0 = __dadd_rn(r0, a[MAXN >= 16 ? 8 : 0]);  r1 = __dadd_rn(r1, a[MAXN >= 16 ? 9 : 0]);
            r2 = __dadd_rn(r2, a[MAXN >= 16 ? 10 : 0]); r3 = __dadd_rn(r3, a[MAXN >= 16 ? 11 : 0]);
            r4 = __dadd_rn(r4, a[MAXN >= 16 ? 12 : 0]); r5 = __dadd_rn(r5, a[MAXN >= 16 ? 13 : 0]);
            r6 = __dadd_rn(r6, a[MAXN >= 16 ? 14 : 0]); r7 = __dadd_rn(r7, a[MAXN >= 16 ? 15 : 0]);
        }
        res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                        __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
        if (n < 16) {
#pragma unroll
            for (int j = 8; j < 15; ++j)
                if (j < MAXN && j < n) res = __dadd_rn(res, a[j]);
        }
    }
    return __ddiv_rn(__dadd_rn(0.0, res), (double)n);
}

// sliding window of the last `cap` values, oldest first in a[0..cnt-1].
// Steady state (window full and cap == MAXN): a plain register shift.  While the window fills (or for
// a shorter configured window) every slot is rewritten through selects, so the array stays in
// registers (conditional stores would be merged by the compiler into one dynamically indexed store,
// forcing the window into local memory).
template <int MAXN>
__device__ __forceinline__ void window_push(double (&a)[MAXN], int &cnt, int cap, double v)
{
    if (cnt == MAXN) {  // implies cap == MAXN
#pragma unroll
        for (int j = 0; j < MAXN - 1; ++j) a[j] = a[j + 1];
        a[MAXN - 1] = v;
        return;
    }
    const bool full = cnt >= cap;
#pragma unroll
    for (int j = 0; j < MAXN; ++j) {
        const double next = (j + 1 < MAXN) ? a[j + 1 < MAXN ? j + 1 : j] : v;
        const double when_full = (j == cap - 1) ? v : ((j < cap - 1) ? next : a[j]);
        const double when_filling = (j == cnt) ? v : a[j];
        a[j] = full ? when_full : when_filling;
    }
    cnt += full ? 0 : 1;
}

// numpy's pairwise sum of a completely filled 15-slot window (static order): 8 accumulators, tree, tail
__device__ __forceinline__ double np_sum_full(const double (&a)[15])
{
    double res = __dadd_rn(__dadd_rn(__dadd_rn(a[0], a[1]), __dadd_rn(a[2], a[3])),
                           __dadd_rn(__dadd_rn(a[4], a[5]), __dadd_rn(a[6], a[7])));
#pragma unroll
    for (int j = 8; j < 15; ++j) res = __dadd_rn(res, a[j]);
    return res;
}

// np.mean of a completely filled window (n == MAXN), static summation order
template <int MAXN>
__device__ __forceinline__ double np_mean_full(const double (&a)[MAXN])
{
    static_assert(MAXN >= 8 && MAXN <= 16, "window length");
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    if (MAXN == 16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[MAXN == 16 ? 8 + j : j]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    if (MAXN < 16) {
#pragma unroll
        for (int j = 8; j < MAXN; ++j) res = __dadd_rn(res, a[j]);
    }
    return __ddiv_rn(res, (double)MAXN);
}

__device__ __forceinline__ double sigmoid_prob(double beta, double metric, double intercept)
{
    // 1.0 / (1.0 + np.exp(-beta * (metric - intercept)))   cancer_simulation.py:322-323
    const double z = __dmul_rn(-beta, __dsub_rn(metric, intercept));
    return __ddiv_rn(1.0, __dadd_rn(1.0, exp(z)));
}

// V * (1 + rho*log(K/V) - beta_c*C - (alpha*d + beta*d**2) + noise)   cancer_simulation.py:300-302
__device__ __forceinline__ double gompertz_step(const Patient &p, double V, double C, double D, double noise)
{
    double s = __dadd_rn(1.0, __dmul_rn(p.rho, log(__ddiv_rn(p.K, V))));
    s = __dsub_rn(s, __dmul_rn(p.beta_c, C));
    s = __dsub_rn(s, __dadd_rn(__dmul_rn(p.alpha, D), __dmul_rn(p.beta, __dmul_rn(D, D))));
    s = __dadd_rn(s, noise);
    return __dmul_rn(V, s);
}

// u < exp(-V * TUMOUR_CELL_DENSITY) (strict, :346) or u <= ... (non-strict, :551/:759).
// exp underflows to exactly 0 below -745.2; u >= 0, so the transcendental is skipped there.
template <bool STRICT>
__device__ __forceinline__ bool recovery_test(double u, double V, double density)
{
    const double x = __dmul_rn(-V, density);
    if (x < -746.0) return STRICT ? false : (u <= 0.0);
    if (!STRICT && x > -40.0 && x < 0.0) return u <= fm::exp_fast(x);   // K2 / K3: the lean exponential K1 runs here too
    const double e = exp(x);
    return STRICT ? (u < e) : (u <= e);
}

// ---- population statistics accumulated per patient (theta_gram, fused or standalone) ------------
// For one patient (static feature u constant) every entry of Theta^T Theta / Theta^T xdot with
// Theta = [1, x, u, x*u] is u^p times one of five sums, kept per treatment:
//   s[a] = { n, sum x, sum x^2, sum xdot, sum x*xdot }
struct PatientGram {
    double s[4][5];
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int m = 0; m < 5; ++m) s[a][m] = 0.0;
    }
    // one library row theta(x,u) with target xdot, filed under treatment `code`
    __device__ __forceinline__ void add(int code, double x, double xdot)
    {
        const double xx = x * x, xd = x * xdot;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            if (code == a) {
                s[a][0] += 1.0; s[a][1] += x; s[a][2] += xx; s[a][3] += xdot; s[a][4] += xd;
            }
        }
    }
};

// expand one patient's sums of treatment a into the 15 packed statistics (b200i.h layout):
// G00 G01 G02 G03 G11 G12 G13 G22 G23 G33 | b0 b1 b2 b3 | count
__device__ __forceinline__ void expand_gram(const double (&s)[5], double u, double (&g)[B200I_GRAM_PER_TREATMENT])
{
    const double uu = u * u;
    g[0] = s[0];       g[1] = s[1];       g[2] = u * s[0];   g[3] = u * s[1];
    g[4] = s[2];       g[5] = u * s[1];   g[6] = u * s[2];
    g[7] = uu * s[0];  g[8] = uu * s[1];
    g[9] = uu * s[2];
    g[10] = s[3];      g[11] = s[4];      g[12] = u * s[3];  g[13] = u * s[4];
    g[14] = s[0];
}

}  // namespace b200i
