// philox.cuh -- counter-based generator of the simulator's random draws (throughput mode).
//
// The reference draws four (N,T) arrays from numpy's global MT19937 stream before its loop
// (cancer_simulation.py:275-279): noise = 0.01 * randn, recovery / chemo / radio = rand.  A sequential stream cannot
// be sharded, so the device generator is Philox4x32-10 (Salmon et al., SC'11) keyed by the seed and counted by
// (global patient index, column pair, stream): the draws of a patient do not depend on the launch shape or on the
// number of GPUs.  One call yields the 2 x 64 bits of one stream for the two columns of a pair:
//     counter = (patient lo, patient hi, column / 2, stream),  key = (seed lo, seed hi)
//     stream 0: Box-Muller pair -> noise of the even / odd column;  1: recovery;  2: chemo;  3: radio
//     uniform: the top 52 bits of a 64-bit word fill the mantissa of a double in [1,2); u = that - 1  (in [0,1));
//              the recovery stream adds 2^-53 (bin centres, in (0,1)): its smallest value exceeds exp(-40), so the
//              simulator kernel only has to generate it when the recovery test can succeed at all (V < 6.9e-8)
//     normal:  r = sqrt(-2 log(2 - m1)), (z_even, z_odd) = r * (cospi, sinpi)(2 * (m2 - 1)),  noise = 0.01 * z
// The test-suite restates this in numpy and pins the block function to the Random123 known-answer vectors.
#pragma once
#include <stdint.h>

namespace b200i {
namespace rng {

struct U4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1)
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
        U4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
        n.w = (uint32_t)p0;
        c = n;
        k0 += W0; k1 += W1;
    }
    return c;
}

// The ten round keys of a seed.  Kernels receive them as a __grid_constant__ parameter: every use is then a
// constant-bank operand of the XOR instead of two uniform-datapath additions per round and call.
struct RoundKeys {
    uint32_t k0[10], k1[10];
};
__host__ __device__ inline RoundKeys round_keys(uint32_t k0, uint32_t k1)
{
    RoundKeys rk;
    for (int r = 0; r < 10; ++r) {
        rk.k0[r] = k0; rk.k1[r] = k1;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return rk;
}
__host__ __device__ __forceinline__ U4 philox4x32_10(U4 c, const RoundKeys &rk)
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
        U4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ rk.k0[r];
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ rk.k1[r];
        n.w = (uint32_t)p0;
        c = n;
    }
    return c;
}

#ifdef __CUDACC__
// double in [1,2) whose mantissa is the top 52 bits of (hi:lo)
__device__ __forceinline__ double mant12(uint32_t lo, uint32_t hi)
{
    return __hiloint2double((int)(0x3FF00000u | (hi >> 12)), (int)((hi << 20) | (lo >> 12)));
}

struct PairKey {
    uint32_t p_lo, p_hi, k0, k1;
    const RoundKeys *rk;   // round keys of (k0, k1), normally the kernel's __grid_constant__ parameter
};

// the two uniforms in [0,1) of stream s (1..3) for column pair tp
__device__ __forceinline__ void uniform_pair(const PairKey &k, uint32_t tp, uint32_t s, double &u_even, double &u_odd)
{
    const U4 r = philox4x32_10(U4{k.p_lo, k.p_hi, tp, s}, *k.rk);
    u_even = __dsub_rn(mant12(r.x, r.y), 1.0);
    u_odd = __dsub_rn(mant12(r.z, r.w), 1.0);
}

// the two recovery uniforms of column pair tp (stream 1): bin centres, in (0,1)
__device__ __forceinline__ void recovery_pair(const PairKey &k, uint32_t tp, double &u_even, double &u_odd)
{
    const U4 r = philox4x32_10(U4{k.p_lo, k.p_hi, tp, 1u}, *k.rk);
    u_even = __dadd_rn(__dsub_rn(mant12(r.x, r.y), 1.0), 0x1p-53);
    u_odd = __dadd_rn(__dsub_rn(mant12(r.z, r.w), 1.0), 0x1p-53);
}

// recovery draw of one column, generated on demand (rare: kept out of line so that the callers' loops stay small)
__device__ __noinline__ double recovery_draw(uint32_t p_lo, uint32_t p_hi, uint32_t k0, uint32_t k1, uint32_t tp, uint32_t odd)
{
    double a, b;
    const RoundKeys rk = round_keys(k0, k1);
    recovery_pair(PairKey{p_lo, p_hi, k0, k1, &rk}, tp, a, b);
    return odd ? b : a;
}
struct LazyRecovery {
    const PairKey *key;
    uint32_t tp, odd;
    __device__ __forceinline__ double get() const { return recovery_draw(key->p_lo, key->p_hi, key->k0, key->k1, tp, odd); }
};

// the two noise terms 0.01 * N(0,1) of column pair tp (stream 0)
__device__ __forceinline__ void noise_pair(const PairKey &k, uint32_t tp, double &z_even, double &z_odd)
{
    const U4 r = philox4x32_10(U4{k.p_lo, k.p_hi, tp, 0u}, *k.rk);
    const double u1 = __dsub_rn(2.0, mant12(r.x, r.y));                  // (0,1]
    const double a2 = __dmul_rn(2.0, __dsub_rn(mant12(r.z, r.w), 1.0));  // [0,2): angle / pi
    const double rad = __dmul_rn(0.01, sqrt(__dmul_rn(-2.0, log(u1))));
    double sn, cs;
    sincospi(a2, &sn, &cs);
    z_even = __dmul_rn(rad, cs);
    z_odd = __dmul_rn(rad, sn);
}
#endif

}  // namespace rng
}  // namespace b200i
