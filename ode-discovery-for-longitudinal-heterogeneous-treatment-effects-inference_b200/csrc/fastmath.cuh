// fastmath.cuh -- lean FP64 log / exp / cbrt / division for the simulator kernels.
//
// ncu (profiles/r1_k1_gen1_variant2_ncu.txt) shows the simulator is instruction-issue bound, not FP64 bound:
// 489 warp instructions per patient-column, only 25 % of them FP64 arithmetic.  CUDA's libdevice
// log/exp/cbrt and IEEE division spend most of that on constant materialisation (UMOV / IMAD.MOV pairs),
// special-case handling and slow paths that can never trigger on this path (arguments are positive, finite
// and normal).  The versions below keep their coefficients in __constant__ memory (an FP64 instruction reads
// a constant-bank operand directly), use MUFU seeds + Newton steps instead of IEEE division, skip the
// special cases, and fuse the division of log(K/V) into the log's own argument reduction.
//
// Accuracy (tests/test_fastmath.py compiles this header for the host and compares with glibc long double):
//   div_fast, rcp_fast : correctly rounded except in rare last-bit cases (< 0.501 ulp)
//   log_ratio(a, b)    : < 0.8 ulp of log(a/b) for |log(a/b)| >= 1 (the path has K/V >= 12)
//   exp_fast           : < 1.1 ulp on [-700, 700]
//   cbrt_fast          : < 0.6 ulp
// i.e. the same class of deviation from numpy's own SIMD log/exp/pow as libdevice's <= 1-2 ulp functions.
//
// Every operation is written with explicit single-rounding helpers (mul/add/sub -> __dmul_rn/... on the
// device) and explicit fma(), so the device and the host build of this header execute the same sequence of
// IEEE operations; only the hardware seeds (MUFU.RCP64H, MUFU.LG2/EX2) differ, and they are refined to
// far below the final rounding.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define B200I_HD __host__ __device__ __forceinline__
#else
#define B200I_HD inline
#endif

namespace b200i {
namespace fm {

#if defined(__CUDA_ARCH__)
#define B200I_CONST __constant__
#else
#define B200I_CONST static const
#endif

// 2*atanh(s) = 2s + s*z*(Lg1 + Lg2 z + ... + Lg7 z^6), z = s^2, |s| <= 0.1716 (fdlibm e_log.c minimax set)
B200I_CONST double kLg[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                             2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                             1.479819860511658591e-01};
B200I_CONST double kLn2Hi = 6.93147180369123816490e-01, kLn2Lo = 1.90821492927058770002e-10;
B200I_CONST double kSqrt2 = 1.4142135623730951;
// 1/n!, n = 2..13
B200I_CONST double kExpC[12] = {5.0e-01, 1.6666666666666666e-01, 4.1666666666666664e-02, 8.3333333333333332e-03,
                                1.3888888888888889e-03, 1.9841269841269841e-04, 2.4801587301587302e-05,
                                2.7557319223985893e-06, 2.7557319223985888e-07, 2.5052108385441720e-08,
                                2.0876756987868100e-09, 1.6059043836821613e-10};
B200I_CONST double kLog2e = 1.4426950408889634074;
// 1/n and n for the window mean while the 15-slot window fills (index = n, 0 unused)
B200I_CONST double kInvN[16] = {0.0, 1.0, 0.5, 1.0 / 3.0, 0.25, 0.2, 1.0 / 6.0, 1.0 / 7.0, 0.125, 1.0 / 9.0, 0.1,
                                1.0 / 11.0, 1.0 / 12.0, 1.0 / 13.0, 1.0 / 14.0, 1.0 / 15.0};

// ---- single-rounding primitives (no FMA contraction on either side) ---------------------------
B200I_HD double mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
B200I_HD double add(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
B200I_HD double sub(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    volatile double r = a - b;
    return r;
#endif
}

B200I_HD double hi_lo_to_double(int hi, unsigned lo)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, (int)lo);
#else
    uint64_t b = ((uint64_t)(uint32_t)hi << 32) | lo;
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
B200I_HD int double_hi(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    return (int)(b >> 32);
#endif
}
B200I_HD unsigned double_lo(double x)
{
#if defined(__CUDA_ARCH__)
    return (unsigned)__double2loint(x);
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    return (unsigned)b;
#endif
}

// The polynomial / reduction constants as a value type: kernels load it once (consts()) and keep the members
// in registers instead of re-reading constant memory for every use (50 LDC/LDCU per column otherwise).
struct FmK {
    double lg[7];
    double ln2hi, ln2lo, sqrt2, log2e;
    double ec[12];
};
B200I_HD FmK consts()
{
    FmK k;
#pragma unroll
    for (int i = 0; i < 7; ++i) k.lg[i] = kLg[i];
    k.ln2hi = kLn2Hi; k.ln2lo = kLn2Lo; k.sqrt2 = kSqrt2; k.log2e = kLog2e;
#pragma unroll
    for (int i = 0; i < 12; ++i) k.ec[i] = kExpC[i];
    return k;
}

// reciprocal seed, relative error <= 2^-22 (MUFU.RCP64H on the device)
B200I_HD double rcp_seed(double x)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
#else
    return (double)(1.0f / (float)x);
#endif
}

// 1/x to ~2^-44: seed + one Newton step.  x normal, finite, non-zero.
B200I_HD double rcp_mid(double x)
{
    const double r = rcp_seed(x);
    return fma(r, fma(-x, r, 1.0), r);
}

// a / b: q0 = a*r, one exact-residual correction -> correctly rounded up to rare last-bit ties
B200I_HD double div_fast(double a, double b)
{
    const double r = rcp_mid(b);
    const double q = mul(a, r);
    return fma(fma(-q, b, a), r, q);
}

// 1 / x
B200I_HD double rcp_fast(double x)
{
    const double r = rcp_mid(x);
    return fma(fma(-r, x, 1.0), r, r);
}

// x / n for a constant n given as (n, fl(1/n)): exact-residual correction, correctly rounded
B200I_HD double div_small(double x, double n, double inv_n)
{
    const double q = mul(x, inv_n);
    return fma(fma(-q, n, x), inv_n, q);
}

// Per-numerator constants of log_ratio(a, .): computed once per patient (a = carrying capacity K).
struct LogNum {
    double m;      // mantissa of a in [1, 2)
    double m_r2;   // m * sqrt(2)
    int ebits;     // biased exponent field of a (hi word & 0x7ff00000)
};
B200I_HD LogNum log_num(double a)
{
    LogNum n;
    const int hi = double_hi(a);
    n.ebits = hi & 0x7ff00000;
    n.m = hi_lo_to_double((hi & 0x000fffff) | 0x3ff00000, double_lo(a));
    n.m_r2 = mul(n.m, kSqrt2);
    return n;
}

// log(a / b) for positive, finite, normal a and b, without forming the quotient:
//   a = A 2^ea, b = B 2^eb, A,B in [1,2); A or B is doubled so that A/B lies in [sqrt(1/2), sqrt(2)];
//   s = (A - B) / (A + B)  (numerator exact by Sterbenz), log(A/B) = 2 atanh(s).
B200I_HD double log_ratio(const FmK &K, const LogNum &na, double b)
{
    const int hb = double_hi(b);
    int k = (na.ebits - (hb & 0x7ff00000)) >> 20;
    const double B = hi_lo_to_double((hb & 0x000fffff) | 0x3ff00000, double_lo(b));
    const bool big = na.m > mul(B, K.sqrt2);  // A/B > sqrt2  -> compare against 2B
    const bool small = na.m_r2 < B;           // A/B < 1/sqrt2 -> use 2A
    const double A2 = hi_lo_to_double(double_hi(na.m) + (small ? 0x00100000 : 0), double_lo(na.m));
    const double B2 = hi_lo_to_double(double_hi(B) + (big ? 0x00100000 : 0), double_lo(B));
    k += (big ? 1 : 0) - (small ? 1 : 0);
    const double s = div_fast(sub(A2, B2), add(A2, B2));
    const double dk = (double)k;
    const double z = mul(s, s);
    const double w = mul(z, z);
    const double t1 = mul(w, fma(w, fma(w, K.lg[5], K.lg[3]), K.lg[1]));
    const double t2 = mul(z, fma(w, fma(w, fma(w, K.lg[6], K.lg[4]), K.lg[2]), K.lg[0]));
    const double sR = mul(s, add(t2, t1));
    const double lo = fma(dk, K.ln2lo, sR);
    return fma(dk, K.ln2hi, fma(2.0, s, lo));
}
B200I_HD double log_ratio(double a, double b) { return log_ratio(consts(), log_num(a), b); }

// ---- table-driven log for the projected steps of the treatment-sequence generator (csrc/sim_cf_seq.cuh) ----------
// log(x) = e ln2 + log(c_j) + log1p(r): x = m 2^e, m in [1,2), j = top LOGT_BITS mantissa bits, c_j the interval's
// midpoint, r = m / c_j - 1.  The table holds (invc_j, logc_j) with invc_j = 1/c_j rounded to 13 significant bits and
// logc_j = -log(invc_j), so r = fma(m, invc_j, -1) carries no cancellation error (glibc's construction), |r| < 2^-8
// + 2^-13 and a degree-6 polynomial leaves 3e-18.  Absolute error < 2 ulp of the result for |log| >= 1: the projected
// step multiplies the logarithm by rho ~ 1e-4..3e-2 before it meets a number of order 1, so this is far below the
// last bit of the volume; the factual steps (whose comparisons must match bit for bit) do not use it.
constexpr int LOGT_BITS = 7;
constexpr int LOGT_SIZE = 1 << LOGT_BITS;
struct LogTabEntry {
    double invc, logc;
};
// entry j of the table (filled once per kernel into shared memory / once per process on the host)
B200I_HD LogTabEntry log_table_entry(int j)
{
    const double c = 1.0 + ((double)j + 0.5) / (double)LOGT_SIZE;
    const double inv = 1.0 / c;
    LogTabEntry e;
    e.invc = hi_lo_to_double(double_hi(inv) & 0xffffff00, 0u);   // 12 explicit mantissa bits
    e.logc = -log(e.invc);
    return e;
}
B200I_CONST double kLp[5] = {-0.5, 1.0 / 3.0, -0.25, 0.2, -1.0 / 6.0};   // log1p(r) = r + r^2 (c2 + c3 r + ... + c6 r^4)
struct LogTabK {
    double ln2, c2, c3, c4, c5, c6;
};
B200I_HD LogTabK log_tab_consts()
{
    LogTabK k;
    k.ln2 = 0.693147180559945309417232; k.c2 = kLp[0]; k.c3 = kLp[1]; k.c4 = kLp[2]; k.c5 = kLp[3]; k.c6 = kLp[4];
    return k;
}
B200I_HD int log_tab_index(double x) { return (double_hi(x) >> (20 - LOGT_BITS)) & (LOGT_SIZE - 1); }
// `base` - log(x) for a positive, finite, normal x (base = log of the numerator of a ratio); t = entry log_tab_index(x)
B200I_HD double base_minus_log_entry(const LogTabK &K, const LogTabEntry t, double base, double x)
{
    const int hx = double_hi(x);
    const int e = (hx >> 20) - 1023;
    const double m = hi_lo_to_double((hx & 0x000fffff) | 0x3ff00000, double_lo(x));
    const double r = fma(m, t.invc, -1.0);
    const double r2 = mul(r, r);
    double a = fma(r, K.c3, K.c2);
    double b = fma(r, K.c5, K.c4);
    b = fma(r2, K.c6, b);
    a = fma(r2, b, a);
    const double p = fma(r2, a, r);                   // log1p(r)
    const double u = sub(fma(-(double)e, K.ln2, base), t.logc);
    return sub(u, p);
}
B200I_HD double base_minus_log_tab(const LogTabK &K, const LogTabEntry *__restrict__ tab, double base, double x)
{
    return base_minus_log_entry(K, tab[log_tab_index(x)], base, x);
}

// exp(x) for |x| <= 700 (callers route anything else to the library function)
B200I_HD double exp_fast(const FmK &K, double x)
{
    const double shift = 6755399441055744.0;   // 1.5 * 2^52
    const double kd_s = fma(x, K.log2e, shift);
    const int k = (int)double_lo(kd_s);        // round-to-nearest integer sits in the low word
    const double kd = sub(kd_s, shift);
    double r = fma(-kd, K.ln2hi, x);
    r = fma(-kd, K.ln2lo, r);
    // exp(r) = 1 + r + r^2 (E(r^2) + r O(r^2)): two Horner chains of half the length
    const double r2 = mul(r, r);
    double pe = fma(r2, K.ec[10], K.ec[8]);   // 1/12!, 1/10!
    double po = fma(r2, K.ec[11], K.ec[9]);   // 1/13!, 1/11!
    pe = fma(r2, pe, K.ec[6]);  po = fma(r2, po, K.ec[7]);
    pe = fma(r2, pe, K.ec[4]);  po = fma(r2, po, K.ec[5]);
    pe = fma(r2, pe, K.ec[2]);  po = fma(r2, po, K.ec[3]);
    pe = fma(r2, pe, K.ec[0]);  po = fma(r2, po, K.ec[1]);
    const double q = fma(r, po, pe);
    const double p = fma(r2, q, r);             // exp(r) - 1, |p| < 0.42
    const double e = add(1.0, p);
    return hi_lo_to_double(double_hi(e) + (int)((unsigned)k << 20), double_lo(e));   // e * 2^k, result stays normal
}
B200I_HD double exp_fast(double x) { return exp_fast(consts(), x); }

// cube root of a positive, finite x that is representable as a normal float
B200I_HD double cbrt_fast(double x)
{
    const float xf = (float)x;
#if defined(__CUDA_ARCH__)
    float lg, y0f;   // MUFU.LG2 / MUFU.EX2 without the denormal fix-up sequences, ~2^-21
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(xf));
    lg *= 0.33333334f;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y0f) : "f"(lg));
    const double y0 = (double)y0f;
#else
    const double y0 = (double)cbrtf(xf);
#endif
    const double y2 = mul(y0, y0);                 // exact: y0 has 24 significant bits
    const double e = fma(-y2, y0, x);              // x - y0^3, one rounding
    const double den = fma(-2.0, e, mul(3.0, x));  // 2 y0^3 + x
    // Halley: y0 + y0 e / (2 y0^3 + x); the correction is 2^-21 relative, so 1/den to 2^-44 is ample
    return fma(mul(y0, e), rcp_mid(den), y0);
}

}  // namespace fm
}  // namespace b200i
