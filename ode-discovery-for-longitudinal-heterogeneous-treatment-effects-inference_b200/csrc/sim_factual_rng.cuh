// sim_factual_rng.cuh -- K1L: simulate_factual with the random draws generated in registers (throughput mode).
// Included by sim_factual.cu after sim_factual_ws.cuh (reuses its column arithmetic ws_body / ws_treat).
//
// The reference draws four (N,T) float64 arrays before its loop (cancer_simulation.py:275-279) and K1 reads them
// from HBM: 1.9 of the 6.3 GB a 1M-patient launch moves, and 2 GB per step over PCIe when the cohort starts on the
// host.  SURVEY.md 8(d) names the lean variant: device generator, outputs = volume (float64) + one treatment-code
// byte per step + sequence length.  The kernels of this file are that variant:
//   * thread = patient, warp = 32 consecutive patients, no CTA-wide synchronisation in the main loop;
//   * draws: Philox4x32-10 counted by (global patient, column pair, stream) -- philox.cuh -- so a patient's draws do
//     not depend on the launch shape, the shard or the number of GPUs; b200i_philox_draws exports the same numbers
//     as the four (N,T) arrays, and b200i_sim_factual on those arrays reproduces these kernels bit for bit (tested);
//   * column arithmetic: ws_body / ws_col = the lean three-chain column of the tiled kernel (fastmath.cuh); tiles
//     outside its preconditions take the generic column function with the same draws;
//   * volume: 16-column x 32-patient SWIZZLE_128B tiles in shared memory, stored by TMA (cp.async.bulk.tensor);
//   * treatment codes chemo + 2*radio: one byte per step, staged per tile in shared memory (odd word pitch) and
//     written with coalesced 16-byte stores; layout = what b200i_theta_gram_codes reads;
//   * STATS 2: the six per-patient moment sums of get_scaling_params (as b200i_sim_factual_side);
//     STATS 1: the population statistics of K4 accumulated on the fly (Gram + moments), reduced in a fixed order.
// Two generations (identical bits; `variant` of b200i_sim_factual_rng): sim_factual_rng_kernel inlines the generator
// into a four-column loop body (12 warps / SM; bound by instruction fetch), sim_factual_rng2_kernel (the default,
// further down) splits generator and simulator into two short loops (16 warps / SM; bound by instruction issue).
// HBM traffic: 80 B read + T*8 + ceil16(T) + 8 (+48) B written per patient = 7 % of the HBM peak at 1.17 ms per
// million patients -- these kernels are bound by the SM front end and the FP64 pipe, not by memory.
#pragma once
#include "philox.cuh"

namespace b200i {

constexpr int RNG_WARPS = 4;
constexpr int RNG_VOL_BYTES = 2 * 32 * 128;   // two tiles of 32 patients x 16 columns per warp

__host__ __device__ inline int rng_code_copy_words(int T) { return ((T + 15) / 16) * 4; }     // words written per row
__host__ __device__ inline int rng_code_words(int T) { return rng_code_copy_words(T) + 1; }    // odd pitch in smem
__host__ __device__ inline int rng_warp_bytes(int T) { return RNG_VOL_BYTES + ((32 * rng_code_words(T) * 4 + 1023) & ~1023); }

// one 16-column box of a tile on the generic column function (library log / exp / cbrt)
template <bool STAT>
__device__ __noinline__ void rng_slow_box(uint8_t *buf, uint8_t *crow, int lane, int m, int T, const SimC &c, WsSlow *st,
                                          const rng::PairKey &key)
{
    for (int cidx = 0; cidx < 16; cidx += 2) {
        const int t = 16 * m + cidx;
        double nz[2], ur[2], uc[2], ud[2], v2[2];
        rng::noise_pair(key, (uint32_t)(t >> 1), nz[0], nz[1]);
        rng::recovery_pair(key, (uint32_t)(t >> 1), ur[0], ur[1]);
        rng::uniform_pair(key, (uint32_t)(t >> 1), 2u, uc[0], uc[1]);
        rng::uniform_pair(key, (uint32_t)(t >> 1), 3u, ud[0], ud[1]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            Column o;
            o.V = 0.0;
            if (t + j < T) {
                factual_column<STAT, false>(t + j, T, c, st->p, st->s, nz[j], ur[j], uc[j], ud[j], nullptr, o, st->pg,
                                            st->mom);
                crow[t + j] = (uint8_t)((o.ca != 0.0 ? 1u : 0u) | (o.ra != 0.0 ? 2u : 0u));
            }
            v2[j] = o.V;
        }
        *reinterpret_cast<double2 *>(buf + swz_off<128>((uint32_t)lane, (uint32_t)(cidx >> 1))) = make_double2(v2[0], v2[1]);
    }
}

template <int STATS, int MINB>
__global__ void __launch_bounds__(RNG_WARPS * 32, MINB)
sim_factual_rng_kernel(const __grid_constant__ CUtensorMap vmap, int64_t n, int64_t pstride, int64_t mstride, int T, SimC c,
                       const double *__restrict__ params, const __grid_constant__ rng::RoundKeys rkeys, uint32_t seed_lo,
                        uint32_t seed_hi, int64_t patient_base,
                       uint8_t *__restrict__ codes_out, int64_t code_pitch, double *__restrict__ seq_len_out,
                       double *__restrict__ pmom_out, const double *__restrict__ static_feature, StatsWorkspace *ws)
{
    constexpr bool GRAM = STATS == 1, SIDE = STATS == 2;
    extern __shared__ uint8_t smem_raw[];
    __shared__ double block_acc[GRAM ? RNG_WARPS : 1][STATS_PAD];
    __shared__ unsigned int s_is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t *smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *vt = smem_al + (size_t)warp * rng_warp_bytes(T);          // [2][32 x 128 B], swizzled
    uint32_t *ct = reinterpret_cast<uint32_t *>(vt + RNG_VOL_BYTES);   // [32][code_words]
    const int code_words = rng_code_words(T), copy_words = rng_code_copy_words(T);
    uint32_t *crow = ct + lane * code_words;

    if (GRAM) {
        for (int j = tid; j < RNG_WARPS * STATS_PAD; j += RNG_WARPS * 32) (&block_acc[0][0])[j] = 0.0;
        __syncthreads();
    }
    if (lane == 0) tma_prefetch_desc(&vmap);
    for (int j = 0; j < code_words; ++j) crow[j] = 0u;   // words past the last quad stay zero

    // constants -> registers (as sim_factual_ws)
    WsK k;
    k.f = fm::consts();
    k.sphere = c.sphere; k.inv_sphere = c.inv_sphere; k.decay = c.decay; k.dose = c.chemo_amt; k.death = c.death;
    k.ndensity = -c.density; k.inv15 = fm::kInvN[15];
#pragma unroll
    for (int i = 0; i < 7; ++i) pin(k.f.lg[i]);
#pragma unroll
    for (int i = 0; i < 12; ++i) pin(k.f.ec[i]);
    pin(k.f.ln2hi); pin(k.f.ln2lo); pin(k.f.sqrt2); pin(k.f.log2e);
    pin(k.sphere); pin(k.inv_sphere); pin(k.decay); pin(k.dose); pin(k.death); pin(k.ndensity); pin(k.inv15);

    const uint32_t off0 = (uint32_t)lane * 128u + (((uint32_t)lane & 7u) << 4);   // swizzled 16-byte unit 0 of the row
    const int Tm1 = T - 1;
    const int nboxes = (T + 15) >> 4, nquads = (T + 3) >> 2;
    const double inv_dt = 1.0 / c.fd_dt;
    const int64_t ntiles = (n + 31) / 32;
    int nstored = 0;   // boxes this warp has handed to TMA (buffer parity)

    for (int64_t tile = (int64_t)blockIdx.x * RNG_WARPS + warp; tile < ntiles; tile += (int64_t)gridDim.x * RNG_WARPS) {
        const int64_t patient = tile * 32 + lane;
        const bool exists = patient < n;
        const int64_t pi = exists ? patient : 0;
        const int64_t gp = patient_base + patient;
        const rng::PairKey key{(uint32_t)gp, (uint32_t)((uint64_t)gp >> 32), seed_lo, seed_hi, &rkeys};

        WsPatient p;
        WsState s;
        WsSlow slow;
        double w[18];
#pragma unroll
        for (int j = 0; j < 18; ++j) w[j] = 0.0;
        const double v0 = exists ? __ldg(params + pi) : 1.0;
        const double alpha = __ldg(params + 1 * pstride + pi), beta = __ldg(params + 3 * pstride + pi);
        const double Kcap = __ldg(params + 5 * pstride + pi);
        const double ci = __ldg(params + 6 * pstride + pi), ri = __ldg(params + 7 * pstride + pi);
        const double cb = __ldg(params + 8 * pstride + pi), rb = __ldg(params + 9 * pstride + pi);
        p.rho = __ldg(params + 2 * pstride + pi);
        p.beta_c = __ldg(params + 4 * pstride + pi);
        p.K = fm::log_num(Kcap);
        p.si = ri;
        p.nb = -rb;
        p.rd = __dadd_rn(__dmul_rn(alpha, c.radio_amt), __dmul_rn(beta, __dmul_rn(c.radio_amt, c.radio_amt)));
        s.V = 1.0; s.Cq = s.zq = s.ucp = s.udp = s.S = 0.0; s.flp = 0u; s.alive = exists; s.t_end = 0;
        // fast-path preconditions (see sim_factual_ws): one sigmoid, argument inside exp_fast's domain, normal K / V0
        const double vmax = fmax(v0, c.death);
        const double dmax = 2.02 * cbrt(vmax * c.inv_sphere);
        const double zmax = fabs(rb) * fmax(fabs(ri), fabs(dmax - ri));
        const bool ok = (ci == ri) && (cb == rb) && (zmax <= 700.0) && (Kcap > 1e-300) && (Kcap < 1e300) &&
                        (v0 > 1e-300) && (v0 < 1e300) && (c.window == 15);
        const bool tile_slow = __any_sync(0xffffffffu, exists && !ok) != 0;
        if (tile_slow) {
            slow.p = load_patient(params, pstride, pi);
            state_init(slow.s, exists);
            slow.pg.clear(); slow.mom.clear();
        }
        PatientGram pg;
        Moments mom;
        pg.clear(); mom.clear();
        double gVm1 = 0.0, gVm2 = 0.0;
        unsigned gcm2 = 0u, g_nra = 0u, carry = 0u;
        int g_tlast = -4;

        for (int m = 0; m < nboxes; ++m) {
            uint8_t *buf = vt + (nstored & 1) * 4096;
            if (tile_slow) {
                rng_slow_box<GRAM || SIDE>(buf, reinterpret_cast<uint8_t *>(crow), lane, m, T, c, &slow, key);
            } else {
                auto run_quads = [&](auto fill_tag) {
                    constexpr bool FILL = decltype(fill_tag)::value;
                    int h_hi = (T - 16 * m + 3) >> 2;
                    h_hi = h_hi > 4 ? 4 : h_hi;
#pragma unroll 1
                    for (int h = 0; h < h_hi; ++h) {
                        const int t0 = 16 * m + 4 * h;
                        const uint32_t tp = (uint32_t)(t0 >> 1);
                        double nz[4], uc[4], ud[4];
                        const rng::LazyRecovery ur[4] = {{&key, tp, 0u}, {&key, tp, 1u}, {&key, tp + 1u, 0u}, {&key, tp + 1u, 1u}};
                        rng::noise_pair(key, tp, nz[0], nz[1]);
                        rng::noise_pair(key, tp + 1u, nz[2], nz[3]);
                        rng::uniform_pair(key, tp, 2u, uc[0], uc[1]);
                        rng::uniform_pair(key, tp + 1u, 2u, uc[2], uc[3]);
                        rng::uniform_pair(key, tp, 3u, ud[0], ud[1]);
                        rng::uniform_pair(key, tp + 1u, 3u, ud[2], ud[3]);
                        double oV[4], oC[4], oP[4];
                        unsigned oF[4];
                        ws_body<FILL, 0>(t0, Tm1, k, p, s, w, v0, nz[0], ur[0], uc[0], ud[0], oV[0], oC[0], oP[0], oF[0]);
                        ws_body<FILL, 1>(t0 + 1, Tm1, k, p, s, w, v0, nz[1], ur[1], uc[1], ud[1], oV[1], oC[1], oP[1], oF[1]);
                        ws_body<FILL, 2>(t0 + 2, Tm1, k, p, s, w, v0, nz[2], ur[2], uc[2], ud[2], oV[2], oC[2], oP[2], oF[2]);
                        ws_body<FILL, 3>(t0 + 3, Tm1, k, p, s, w, v0, nz[3], ur[3], uc[3], ud[3], oV[3], oC[3], oP[3], oF[3]);
#pragma unroll
                        for (int j = 0; j < 14; ++j) w[j] = w[j + 4];
                        if (t0 + 3 > s.t_end) {
                            // columns after the last simulated one stay zero; oV[j] belongs to column t0+j, the
                            // treatment outputs to column t0+j-1
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (t0 + j > s.t_end) oV[j] = 0.0;
                                if (t0 + j - 1 > s.t_end) { oC[j] = 0.0; oF[j] = 0u; }
                            }
                        }
                        if (GRAM) {
                            const unsigned c0 = oF[0] & 3u, c1 = oF[1] & 3u, c2 = oF[2] & 3u, c3 = oF[3] & 3u;
                            const int te = s.t_end;
                            ws_gram_sample(pg, t0 >= 2 && t0 - 2 <= te, t0 - 2 == te || c0 != gcm2, gVm2, gVm1, gcm2,
                                           c.fd_dt, inv_dt);
                            ws_gram_sample(pg, t0 >= 1 && t0 - 1 <= te, t0 - 1 == te || c1 != c0, gVm1, oV[0], c0,
                                           c.fd_dt, inv_dt);
                            ws_gram_sample(pg, t0 <= te, t0 == te || c2 != c1, oV[0], oV[1], c1, c.fd_dt, inv_dt);
                            ws_gram_sample(pg, t0 + 1 <= te, t0 + 1 == te || c3 != c2, oV[1], oV[2], c2, c.fd_dt, inv_dt);
                            gVm2 = oV[2]; gVm1 = oV[3]; gcm2 = c3; g_tlast = t0;
                        }
                        if (GRAM || SIDE) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                mom.sv += oV[j]; mom.svv += oV[j] * oV[j];
                                mom.sc += oC[j]; mom.scc += oC[j] * oC[j];
                                g_nra += (oF[j] >> 1) & 1u;
                            }
                        }
                        const uint32_t offA = off0 ^ ((uint32_t)h << 5), offB = offA ^ 16u;
                        *reinterpret_cast<double2 *>(buf + offA) = make_double2(oV[0], oV[1]);
                        *reinterpret_cast<double2 *>(buf + offB) = make_double2(oV[2], oV[3]);
                        // code word of the previous quad: its last column's treatment is known only now
                        if (t0 > 0) crow[(t0 >> 2) - 1] = carry | ((oF[0] & 3u) << 24);
                        carry = (oF[1] & 3u) | ((oF[2] & 3u) << 8) | ((oF[3] & 3u) << 16);
                    }
                };
                if (m == 0) run_quads(std::true_type{}); else run_quads(std::false_type{});
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&vmap, 16 * m, (int)(tile * 32), buf);
                tma_store_commit();
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the other buffer is free again
            }
            ++nstored;
            __syncwarp();
        }
        if (!tile_slow) crow[nquads - 1] = carry;   // column 4*nquads-1 >= T-1 is never simulated: its code is 0
        const int t_end = tile_slow ? slow.s.t_end : s.t_end;
        if (exists) seq_len_out[patient] = (double)(t_end + 1);
        if (SIDE && exists) {
            const Moments &mm = tile_slow ? slow.mom : mom;
            const double sd = tile_slow ? slow.mom.sd : c.radio_amt * (double)g_nra;
            const double sdd = tile_slow ? slow.mom.sdd : c.radio_amt * c.radio_amt * (double)g_nra;
            pmom_out[0 * mstride + patient] = mm.sv;  pmom_out[1 * mstride + patient] = mm.svv;
            pmom_out[2 * mstride + patient] = mm.sc;  pmom_out[3 * mstride + patient] = mm.scc;
            pmom_out[4 * mstride + patient] = sd;     pmom_out[5 * mstride + patient] = sdd;
        }
        if (GRAM) {
            const double u = exists ? __ldg(static_feature + patient) : 0.0;
            if (tile_slow) {
                factual_finish<true>(c, slow.s, slow.pg);
                fold_patient_stats(block_acc[warp], lane, slow.pg, slow.mom, u, exists, t_end + 1);
            } else {
                ws_gram_sample(pg, g_tlast + 2 <= s.t_end, true, gVm2, gVm1, gcm2, c.fd_dt, inv_dt);
                mom.sd = c.radio_amt * (double)g_nra;
                mom.sdd = c.radio_amt * c.radio_amt * (double)g_nra;
                fold_patient_stats(block_acc[warp], lane, pg, mom, u, exists, t_end + 1);
            }
        }
        // the tile's code bytes: 16 per store, consecutive lanes -> consecutive 16 bytes of a row
        __syncwarp();
        if (codes_out != nullptr) {
            const int units = copy_words >> 2;
            const int rows = (n - tile * 32 < 32) ? (int)(n - tile * 32) : 32;
            for (int e = lane; e < rows * units; e += 32) {
                const int r = e / units, q = e - r * units;
                const uint32_t *src = ct + r * code_words + 4 * q;
                const uint4 v = make_uint4(src[0], src[1], src[2], src[3]);
                *reinterpret_cast<uint4 *>(codes_out + (tile * 32 + r) * code_pitch + 16 * q) = v;
            }
        }
        __syncwarp();
    }
    if (lane == 0) tma_store_wait_all();
    if (GRAM) stats_block_finish(block_acc, RNG_WARPS, ws, &s_is_last);
}

// ------------------------------------------------------------------------------------------------
// Second generation: phased, one column per loop body.
// ncu on the first kernel (profiles/r1_k1l_gen1_ncu.txt): the warps wait for INSTRUCTIONS, not data -- stall
// no_instruction 3.3 per issued instruction, issue slots 39 % busy.  Its loop body (four unrolled columns + eight
// inlined Philox calls + two Box-Muller transforms) is 1700 instructions = 27 KB of SASS against a 6 KB L0 / 32 KB
// L1.5 instruction cache shared by 12 warps at different program counters.  Here the work of eight columns is split
// into two short loops that each fit the L0 cache:
//   G  (per column pair, ~350 instructions): Philox + Box-Muller -> noise, chemo and radio draws into a small
//      per-warp scratch (72-byte pitch: conflict-free 8-byte accesses);
//   S  (per column pair): ws_col = the column arithmetic of ws_body, two unrolled columns with static window
//      offsets, the 17-slot diameter window shifted by two once per pair.
// Same draws, same arithmetic, same outputs as the first kernel, bit for bit.
// ------------------------------------------------------------------------------------------------
constexpr int RNG2_SCRATCH_PITCH = 72;                               // bytes per patient row, 8 columns + pad
constexpr int RNG2_SCRATCH_BYTES = 3 * 32 * RNG2_SCRATCH_PITCH;      // noise, chemo and radio draws of 8 columns
constexpr int RNG2_VOL_BYTES = 32 * 128;                             // one tile of 32 patients x 16 columns per warp

__host__ __device__ inline int rng2_warp_bytes(int T)
{
    return (RNG2_VOL_BYTES + RNG2_SCRATCH_BYTES + 32 * rng_code_words(T) * 4 + 1023) & ~1023;
}

// one column: volume of column t, treatment of column t-1, sigmoid argument of column t (see ws_body).
// J = position in the unrolled column pair.  The diameter window lives in w[17]: before the pair w[0..14] are the 15
// most recent cube roots (oldest first); column J appends its own at w[15 + J] and reads w[1 + J .. 15 + J]; the
// caller shifts the file by two once per pair (7.5 register-pair moves per column).
template <int J, class UR>
__device__ __forceinline__ void ws_col(int t, int Tm1, const WsK &k, const WsPatient &p, WsState &s, double (&w)[17],
                                       double v0, double nz, const UR &ur_in, double uc, double ud, double &oV,
                                       double &oC, unsigned &oF)
{
    if (J == 0 && t == 0) {
        oV = v0; oC = 0.0; oF = 0u;
        s.V = v0;
        return;
    }
    double pr, C1;
    bool ra, ca;
    if (J == 1 && t == 1) {
        pr = 0.0; C1 = 0.0; ra = ca = false;
    } else {
        ws_treat(k, s, pr, ra, ca, C1);
    }
    const double cn = fm::cbrt_fast(fm::div_small(s.V, k.sphere, k.inv_sphere));
    w[15 + J] = cn;
    double mean;
    if (t <= 15) {   // the window fills: numpy's pairwise sum is a running sum plus one 8-leaf tree at t = 8
        s.S = (J == 0 && t == 8) ? ws_tree8(w[8 + J], w[9 + J], w[10 + J], w[11 + J], w[12 + J], w[13 + J], w[14 + J], w[15 + J])
                                 : __dadd_rn(s.S, cn);
        mean = fm::div_small(s.S, (double)t, fm::kInvN[t & 15]);
    } else {
        double r = ws_tree8(w[1 + J], w[2 + J], w[3 + J], w[4 + J], w[5 + J], w[6 + J], w[7 + J], w[8 + J]);
#pragma unroll
        for (int j = 8; j < 15; ++j) r = __dadd_rn(r, w[1 + J + j]);
        mean = fm::div_small(r, 15.0, k.inv15);
    }
    const double z = __dmul_rn(p.nb, __dsub_rn(__dmul_rn(mean, 2.0), p.si));
    double g1 = __dadd_rn(1.0, __dmul_rn(p.rho, fm::log_ratio(k.f, p.K, s.V)));
    g1 = __dsub_rn(g1, __dmul_rn(p.beta_c, C1));
    g1 = __dsub_rn(g1, ra ? p.rd : 0.0);
    g1 = __dadd_rn(g1, nz);
    double Vn = __dmul_rn(s.V, g1);
    const bool act = s.alive && (t < Tm1);
    const bool death = Vn > k.death;
    Vn = death ? k.death : Vn;
    const double x = __dmul_rn(Vn, k.ndensity);
    bool recov = false;
    if (WsUrLazy<UR>::value) {
        if (act && !(x <= -40.0)) {
            const double ur = ws_ur_value(ur_in);
            recov = (x >= 0.0) ? (x == x) : ((x > -40.0) ? (ur < fm::exp_fast(k.f, x)) : ws_recovery_rare(ur, x));
        }
    } else {
        const double ur = ws_ur_value(ur_in);
        if (act && !(x <= -40.0 && ur >= 1e-17))
            recov = (x >= 0.0) ? (x == x) : ((x > -40.0) ? (ur < fm::exp_fast(k.f, x)) : ws_recovery_rare(ur, x));
    }
    recov = recov && !death;
    Vn = recov ? 0.0 : Vn;
    oV = Vn; oC = C1;
    oF = s.flp | (ca ? 1u : 0u) | (ra ? 2u : 0u);
    s.flp = (death ? 4u : 0u) | (recov ? 8u : 0u);
    s.V = Vn; s.Cq = C1; s.zq = z; s.ucp = uc; s.udp = ud;
    s.t_end = act ? t : s.t_end;
    s.alive = act && !(death || recov);
}

template <int STATS, int MINB>
__global__ void __launch_bounds__(RNG_WARPS * 32, MINB)
sim_factual_rng2_kernel(const __grid_constant__ CUtensorMap vmap, int64_t n, int64_t pstride, int64_t mstride, int T, SimC c,
                        const double *__restrict__ params, const __grid_constant__ rng::RoundKeys rkeys, uint32_t seed_lo,
                        uint32_t seed_hi, int64_t patient_base,
                        uint8_t *__restrict__ codes_out, int64_t code_pitch, double *__restrict__ seq_len_out,
                        double *__restrict__ pmom_out, const double *__restrict__ static_feature, StatsWorkspace *ws)
{
    constexpr bool GRAM = STATS == 1, SIDE = STATS == 2;
    extern __shared__ uint8_t smem_raw[];
    __shared__ double block_acc[GRAM ? RNG_WARPS : 1][STATS_PAD];
    __shared__ unsigned int s_is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t *smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *vt = smem_al + (size_t)warp * rng2_warp_bytes(T);                  // [32 x 128 B], swizzled
    uint8_t *nz_row = vt + RNG2_VOL_BYTES + lane * RNG2_SCRATCH_PITCH;           // this patient's 8 noise terms
    uint8_t *uc_row = nz_row + 32 * RNG2_SCRATCH_PITCH;                          // ... chemo uniforms
    uint8_t *ud_row = uc_row + 32 * RNG2_SCRATCH_PITCH;                          // ... radio uniforms
    uint32_t *ct = reinterpret_cast<uint32_t *>(vt + RNG2_VOL_BYTES + RNG2_SCRATCH_BYTES);   // [32][code_words]
    const int code_words = rng_code_words(T), copy_words = rng_code_copy_words(T);
    uint8_t *crow = reinterpret_cast<uint8_t *>(ct + lane * code_words);

    if (GRAM) {
        for (int j = tid; j < RNG_WARPS * STATS_PAD; j += RNG_WARPS * 32) (&block_acc[0][0])[j] = 0.0;
        __syncthreads();
    }
    if (lane == 0) tma_prefetch_desc(&vmap);
    for (int j = 0; j < code_words; ++j) reinterpret_cast<uint32_t *>(crow)[j] = 0u;   // bytes >= T-1 stay zero

    WsK k;
    k.f = fm::consts();
    k.sphere = c.sphere; k.inv_sphere = c.inv_sphere; k.decay = c.decay; k.dose = c.chemo_amt; k.death = c.death;
    k.ndensity = -c.density; k.inv15 = fm::kInvN[15];
#pragma unroll
    for (int i = 0; i < 7; ++i) pin(k.f.lg[i]);
#pragma unroll
    for (int i = 0; i < 12; ++i) pin(k.f.ec[i]);
    pin(k.f.ln2hi); pin(k.f.ln2lo); pin(k.f.sqrt2); pin(k.f.log2e);
    pin(k.sphere); pin(k.inv_sphere); pin(k.decay); pin(k.dose); pin(k.death); pin(k.ndensity); pin(k.inv15);

    const uint32_t row_off = (uint32_t)lane * 128u, row_x = (uint32_t)lane & 7u;
    const int Tm1 = T - 1;
    const int nboxes = (T + 15) >> 4;
    const double inv_dt = 1.0 / c.fd_dt;
    const int64_t ntiles = (n + 31) / 32;

    for (int64_t tile = (int64_t)blockIdx.x * RNG_WARPS + warp; tile < ntiles; tile += (int64_t)gridDim.x * RNG_WARPS) {
        const int64_t patient = tile * 32 + lane;
        const bool exists = patient < n;
        const int64_t pi = exists ? patient : 0;
        const int64_t gp = patient_base + patient;
        const rng::PairKey key{(uint32_t)gp, (uint32_t)((uint64_t)gp >> 32), seed_lo, seed_hi, &rkeys};

        WsPatient p;
        WsState s;
        WsSlow slow;
        double w[17];
#pragma unroll
        for (int j = 0; j < 17; ++j) w[j] = 0.0;
        const double v0 = exists ? __ldg(params + pi) : 1.0;
        const double alpha = __ldg(params + 1 * pstride + pi), beta = __ldg(params + 3 * pstride + pi);
        const double Kcap = __ldg(params + 5 * pstride + pi);
        const double ci = __ldg(params + 6 * pstride + pi), ri = __ldg(params + 7 * pstride + pi);
        const double cb = __ldg(params + 8 * pstride + pi), rb = __ldg(params + 9 * pstride + pi);
        p.rho = __ldg(params + 2 * pstride + pi);
        p.beta_c = __ldg(params + 4 * pstride + pi);
        p.K = fm::log_num(Kcap);
        p.si = ri;
        p.nb = -rb;
        p.rd = __dadd_rn(__dmul_rn(alpha, c.radio_amt), __dmul_rn(beta, __dmul_rn(c.radio_amt, c.radio_amt)));
        s.V = 1.0; s.Cq = s.zq = s.ucp = s.udp = s.S = 0.0; s.flp = 0u; s.alive = exists; s.t_end = 0;
        const double vmax = fmax(v0, c.death);
        const double dmax = 2.02 * cbrt(vmax * c.inv_sphere);
        const double zmax = fabs(rb) * fmax(fabs(ri), fabs(dmax - ri));
        const bool ok = (ci == ri) && (cb == rb) && (zmax <= 700.0) && (Kcap > 1e-300) && (Kcap < 1e300) &&
                        (v0 > 1e-300) && (v0 < 1e300) && (c.window == 15);
        const bool tile_slow = __any_sync(0xffffffffu, exists && !ok) != 0;
        if (tile_slow) {
            slow.p = load_patient(params, pstride, pi);
            state_init(slow.s, exists);
            slow.pg.clear(); slow.mom.clear();
        }
        PatientGram pg;
        Moments mom;
        pg.clear(); mom.clear();
        double gVm1 = 0.0, gVm2 = 0.0;   // V[t-1], V[t-2]
        unsigned gcm2 = 0u, g_nra = 0u;  // treatment of column t-2; radio applications so far

        for (int m = 0; m < nboxes; ++m) {
            uint8_t *buf = vt;
            if (tile_slow) {
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
                rng_slow_box<GRAM || SIDE>(buf, crow, lane, m, T, c, &slow, key);
            } else {
                for (int half = 0; half < 2; ++half) {
                    const int tc0 = 16 * m + 8 * half;
                    if (tc0 >= T) break;
                    // ---- G: draws of the eight columns ----
#pragma unroll 1
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t tp = (uint32_t)(tc0 >> 1) + (uint32_t)q;
                        double a, b;
                        rng::noise_pair(key, tp, a, b);
                        *reinterpret_cast<double *>(nz_row + 16 * q) = a;
                        *reinterpret_cast<double *>(nz_row + 16 * q + 8) = b;
                        rng::uniform_pair(key, tp, 2u, a, b);
                        *reinterpret_cast<double *>(uc_row + 16 * q) = a;
                        *reinterpret_cast<double *>(uc_row + 16 * q + 8) = b;
                        rng::uniform_pair(key, tp, 3u, a, b);
                        *reinterpret_cast<double *>(ud_row + 16 * q) = a;
                        *reinterpret_cast<double *>(ud_row + 16 * q + 8) = b;
                    }
                    // ---- S: the eight columns; the previous box has left the volume tile by now ----
                    if (half == 0) {
                        if (lane == 0) tma_store_wait_read();
                        __syncwarp();
                    }
                    int cend = T - tc0;
                    cend = cend > 8 ? 8 : cend;
#pragma unroll 1
                    for (int cc = 0; cc < cend; cc += 2) {   // T is even: whole pairs
                        double *pv = reinterpret_cast<double *>(buf + row_off + ((((uint32_t)(4 * half + (cc >> 1))) ^ row_x) << 4));
                        auto column = [&](auto jtag) {
                            constexpr int J = decltype(jtag)::value;
                            const int t = tc0 + cc + J;
                            const double nz = *reinterpret_cast<const double *>(nz_row + 8 * (cc + J));
                            const double uc = *reinterpret_cast<const double *>(uc_row + 8 * (cc + J));
                            const double ud = *reinterpret_cast<const double *>(ud_row + 8 * (cc + J));
                            const rng::LazyRecovery ur{&key, (uint32_t)(t >> 1), (uint32_t)J};
                            double oV, oC;
                            unsigned oF;
                            ws_col<J>(t, Tm1, k, p, s, w, v0, nz, ur, uc, ud, oV, oC, oF);
                            // columns after the last simulated one stay zero; oV belongs to column t, the treatment
                            // outputs to column t-1
                            oV = (t > s.t_end) ? 0.0 : oV;
                            const bool dead_prev = t - 1 > s.t_end;
                            oC = dead_prev ? 0.0 : oC;
                            oF = dead_prev ? 0u : oF;
                            const unsigned c1 = oF & 3u;
                            if (GRAM) {
                                // regression sample k = t-2 is complete: x[k+1] and the treatment of column k+1 are known
                                ws_gram_sample(pg, t >= 2 && t - 2 <= s.t_end, t - 2 == s.t_end || c1 != gcm2, gVm2, gVm1,
                                               gcm2, c.fd_dt, inv_dt);
                                gVm2 = gVm1; gVm1 = oV; gcm2 = c1;
                            }
                            if (GRAM || SIDE) {
                                mom.sv += oV; mom.svv += oV * oV;
                                mom.sc += oC; mom.scc += oC * oC;
                                g_nra += (oF >> 1) & 1u;
                            }
                            pv[J] = oV;
                            if (t > 0) crow[t - 1] = (uint8_t)c1;
                        };
                        column(std::integral_constant<int, 0>{});
                        column(std::integral_constant<int, 1>{});
#pragma unroll
                        for (int j = 0; j < 15; ++j) w[j] = w[j + 2];
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&vmap, 16 * m, (int)(tile * 32), buf);
                tma_store_commit();
            }
        }
        const int t_end = tile_slow ? slow.s.t_end : s.t_end;
        if (exists) seq_len_out[patient] = (double)(t_end + 1);
        if (SIDE && exists) {
            const Moments &mm = tile_slow ? slow.mom : mom;
            const double sd = tile_slow ? slow.mom.sd : c.radio_amt * (double)g_nra;
            const double sdd = tile_slow ? slow.mom.sdd : c.radio_amt * c.radio_amt * (double)g_nra;
            pmom_out[0 * mstride + patient] = mm.sv;  pmom_out[1 * mstride + patient] = mm.svv;
            pmom_out[2 * mstride + patient] = mm.sc;  pmom_out[3 * mstride + patient] = mm.scc;
            pmom_out[4 * mstride + patient] = sd;     pmom_out[5 * mstride + patient] = sdd;
        }
        if (GRAM) {
            const double u = exists ? __ldg(static_feature + patient) : 0.0;
            if (tile_slow) {
                factual_finish<true>(c, slow.s, slow.pg);
                fold_patient_stats(block_acc[warp], lane, slow.pg, slow.mom, u, exists, t_end + 1);
            } else {
                // the last sample k = T-2 (x[T-1] = 0 is the never-simulated column)
                ws_gram_sample(pg, T - 2 <= s.t_end, true, gVm2, gVm1, gcm2, c.fd_dt, inv_dt);
                mom.sd = c.radio_amt * (double)g_nra;
                mom.sdd = c.radio_amt * c.radio_amt * (double)g_nra;
                fold_patient_stats(block_acc[warp], lane, pg, mom, u, exists, t_end + 1);
            }
        }
        __syncwarp();
        if (codes_out != nullptr) {
            const int units = copy_words >> 2;
            const int rows = (n - tile * 32 < 32) ? (int)(n - tile * 32) : 32;
            for (int e = lane; e < rows * units; e += 32) {
                const int r = e / units, q = e - r * units;
                const uint32_t *src = ct + r * code_words + 4 * q;
                const uint4 v = make_uint4(src[0], src[1], src[2], src[3]);
                *reinterpret_cast<uint4 *>(codes_out + (tile * 32 + r) * code_pitch + 16 * q) = v;
            }
        }
        __syncwarp();
    }
    if (lane == 0) tma_store_wait_all();
    if (GRAM) stats_block_finish(block_acc, RNG_WARPS, ws, &s_is_last);
}

// the generator's draws as the four (N,T) arrays of the reference contract (noise already multiplied by 0.01):
// thread = (patient, column pair), consecutive lanes -> consecutive 16 bytes of a row
__global__ void __launch_bounds__(256)
philox_draws_kernel(int64_t n, int T, int64_t pitch, const __grid_constant__ rng::RoundKeys rkeys, uint32_t seed_lo,
                    uint32_t seed_hi, int64_t patient_base,
                    double *__restrict__ noise, double *__restrict__ rec, double *__restrict__ chemo,
                    double *__restrict__ radio)
{
    const int half = T >> 1;
    const int64_t total = n * half;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / half;
        const int tp = (int)(e - i * half);
        const int64_t gp = patient_base + i;
        const rng::PairKey key{(uint32_t)gp, (uint32_t)((uint64_t)gp >> 32), seed_lo, seed_hi, &rkeys};
        double a, b;
        const int64_t o = i * pitch + 2 * tp;
        rng::noise_pair(key, (uint32_t)tp, a, b);
        *reinterpret_cast<double2 *>(noise + o) = make_double2(a, b);
        rng::recovery_pair(key, (uint32_t)tp, a, b);
        *reinterpret_cast<double2 *>(rec + o) = make_double2(a, b);
        rng::uniform_pair(key, (uint32_t)tp, 2u, a, b);
        *reinterpret_cast<double2 *>(chemo + o) = make_double2(a, b);
        rng::uniform_pair(key, (uint32_t)tp, 3u, a, b);
        *reinterpret_cast<double2 *>(radio + o) = make_double2(a, b);
    }
}

template <int STATS, int MINB, int GEN = 2>
static int launch_rng(const CUtensorMap &vmap, int64_t n, int64_t pstride, int64_t mstride, int T, const SimC &c,
                      const double *params, uint64_t seed,
                      int64_t patient_base, uint8_t *codes_out, int64_t code_pitch, double *seq_len, double *pmom,
                      const double *static_feature, StatsWorkspace *ws, cudaStream_t st)
{
    void (*kern)(const CUtensorMap, int64_t, int64_t, int64_t, int, SimC, const double *, const rng::RoundKeys, uint32_t, uint32_t,
                 int64_t, uint8_t *,
                 int64_t, double *, double *, const double *, StatsWorkspace *);
    if constexpr (GEN == 2) kern = sim_factual_rng2_kernel<STATS, MINB>;
    else kern = sim_factual_rng_kernel<STATS, MINB>;
    const int smem = RNG_WARPS * (GEN == 2 ? rng2_warp_bytes(T) : rng_warp_bytes(T)) + 1024;
    B200I_REQUIRE(smem <= 227 * 1024, B200I_E_UNSUPPORTED, "sim_factual_rng: T=%d does not fit in shared memory", T);
    // attribute + occupancy once per (device, shared-memory size): a chunked pipeline launches this 16x per 2 ms step
    int per_sm = 0;
    {
        int rc = ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem, RNG_WARPS * 32, &per_sm);
        if (rc) return rc;
    }
    B200I_REQUIRE(per_sm >= 1, B200I_E_UNSUPPORTED, "sim_factual_rng: kernel does not fit on an SM (T=%d)", T);
    const int64_t ntiles = (n + 31) / 32;
    int64_t grid = (int64_t)num_sms() * per_sm;
    const int64_t need = (ntiles + RNG_WARPS - 1) / RNG_WARPS;
    if (grid > need) grid = need;
    if (STATS == 1 && grid > STATS_MAX_BLOCKS) grid = STATS_MAX_BLOCKS;
    kern<<<(unsigned)grid, RNG_WARPS * 32, smem, st>>>(vmap, n, pstride, mstride, T, c, params,
                                                       rng::round_keys((uint32_t)seed, (uint32_t)(seed >> 32)), (uint32_t)seed, (uint32_t)(seed >> 32),
                                                       patient_base, codes_out, code_pitch, seq_len, pmom,
                                                       static_feature, ws);
    return check_cuda(cudaGetLastError(), "sim_factual_rng launch");
}

}  // namespace b200i
