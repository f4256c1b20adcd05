// sim_cf_seq.cuh -- K3, second generation: simulate_counterfactuals_treatment_seq (cancer_simulation.py:635-760,
// 'sliding_treatment') as two kernels per dependency level.
//
// The first generation (cf_treatment_seq_kernel in sim_cf.cu, still used for patient 0, whose window reads the row
// it is writing itself) ran one thread per patient through the factual step AND the 2H projected sequences of every
// step: 86 KB of unrolled SASS, 164 registers, 12 warps per SM, FP64 pipe 31 % busy, 63 % of the issued instructions
// not FP64 (profiles/r1_k3_treatment_seq_ncu.txt).  The projections of a step depend on the factual trajectory only
// through (F[t+1], C[t]), so they are independent across (patient, t):
//
//   phase A  cf_seq_factual_kernel   one thread per patient: cross-row window -> treatment assignment -> factual
//                                    step (:655-705); writes F, the code bytes, the chemo concentration C[t] and
//                                    n_steps.  ~1/8 of the arithmetic.
//   phase B  cf_seq_project_kernel   CTA = 32 patients (lane = patient), work item = (t, s): the two sliding options
//                                    that start at projected step s (chemo at s / radio at s, :575-580) plus the s
//                                    untreated steps before them -- 10 - s Gompertz steps, the two option chains
//                                    independent of each other.  Ten warps pair the items of four consecutive t as
//                                    (s=0, s=4), (s=1, s=3), (s=2, s=2): 16 steps per warp and chunk, one 100-instruction
//                                    loop body, ~70 registers.  Every item writes its own two rows of the (2H, H)
//                                    block into a shared-memory stage; per chunk and patient one bulk copy
//                                    (cp.async.bulk shared -> global, SASS UBLKCP) stores 4 x 400 contiguous bytes.
//
// Projected step (:739-743):  V' = V * (1 + rho * log(K / (V + 1e-07) + 1e-07) - beta_c * C - (alpha d + beta d^2) + e).
// log(q + eps) with q = K / den is evaluated as log_ratio(K, den) + eps * den / K: the quotient is never formed and
// the neglected term (eps / q)^2 / 2 is below 4e-17 for q >= 12 (V below the death threshold) and below 3e-15 for any
// q >= 1.4 -- inside the 1e-9 tolerance of the parity tests by six orders of magnitude.  A non-positive or non-finite
// denominator (negative projected volume: the reference's log yields the NaN that drops the option, :745-746) takes
// the library path with the reference's exact expression, so the validity masks follow the reference.
#pragma once

namespace b200i {

// ---- phase B ------------------------------------------------------------------------------------------------------
constexpr int PB_TT = 4;                       // factual steps per chunk
constexpr int PB_CWARPS = 10;                  // compute warps: (s=0,s=4) x4, (s=1,s=3) x4, (s=2,s=2) x2
constexpr int PB_THREADS = (PB_CWARPS + 1) * 32;   // + one load / store warp
constexpr int PB_H = 5;
constexpr int PB_BLK = 2 * PB_H * PB_H;        // doubles per (patient, t)
constexpr int PB_PSTRIDE = PB_TT * PB_BLK + 2; // doubles per patient in the stage: rows stay 16-byte aligned
constexpr int PB_STAGE_BYTES = 32 * PB_PSTRIDE * 8;
constexpr int PB_IN_DOUBLES = 2 * PB_TT + PB_TT + PB_H - 1;   // per patient and chunk: F[4], C[4], noise[8]
constexpr int PB_IN_STRIDE = PB_IN_DOUBLES + 1;                // odd: all lanes read the same field without bank conflicts
constexpr int PB_IN_BYTES = 32 * PB_IN_STRIDE * 8;
constexpr int PB_VALID_BYTES = 32 * PB_TT * PB_H;             // one byte per (patient, t, s): bit 0 chemo, bit 1 radio option
constexpr int PB_TAB_BYTES = fm::LOGT_SIZE * 16;
// two stage buffers, two input buffers, two validity buffers, the log table, six mbarriers
constexpr int PB_OFF_IN = 2 * PB_STAGE_BYTES;
constexpr int PB_OFF_TAB = PB_OFF_IN + 2 * PB_IN_BYTES;
constexpr int PB_OFF_VALID = PB_OFF_TAB + PB_TAB_BYTES;
constexpr int PB_OFF_BARS = PB_OFF_VALID + 2 * PB_VALID_BYTES;
constexpr int PB_SMEM_BYTES = PB_OFF_BARS + 64;
static_assert(2 * (PB_SMEM_BYTES + 1024) <= 228 * 1024, "two CTAs per SM");

// scalar constants of the projected step as ONE __grid_constant__ kernel parameter: every member is a constant-bank
// operand of the FP64 instruction that uses it (no UMOV / IMAD.MOV pairs to materialise 64-bit immediates)
struct ProjK {
    fm::LogTabK lt;
    double eps, big, decay, chemo_amt;
};

struct ProjPatient {
    double rho, neg_beta_c, K, td, eps_over_K, logK;
};

__device__ __forceinline__ ProjPatient proj_patient(const ProjK &k, double radio_amt, const double *__restrict__ params,
                                                    int64_t n, int64_t i)
{
    ProjPatient p;
    p.rho = __ldg(params + 2 * n + i);
    p.neg_beta_c = -__ldg(params + 4 * n + i);
    p.K = __ldg(params + 5 * n + i);
    const double alpha = __ldg(params + 1 * n + i), beta = __ldg(params + 3 * n + i);
    p.td = __dadd_rn(__dmul_rn(alpha, radio_amt), __dmul_rn(beta, __dmul_rn(radio_amt, radio_amt)));
    p.eps_over_K = __ddiv_rn(k.eps, p.K);
    p.logK = log(p.K);
    return p;
}

// One projected step (:739-743) in two halves, so that the two option chains of an item share one basic block and
// interleave (a validity branch per chain would serialise them).  The logarithm is multiplied by rho (1e-4 .. 3e-2)
// before it meets a number of order one, so the table-driven log (< 6e-15 absolute) and the fused multiply-adds
// below stay far inside the last bit of the step's own rounding class.
// Tab: j -> table entry (shared-memory table in phase B; computed on the fly for patient 0's own first row in phase A)
struct SmemLogTab {
    const fm::LogTabEntry *tab;
    __device__ __forceinline__ fm::LogTabEntry operator()(int j) const { return tab[j]; }
};
struct ComputedLogTab {
    __device__ __forceinline__ fm::LogTabEntry operator()(int j) const { return fm::log_table_entry(j); }
};
template <typename Tab>
__device__ __forceinline__ double proj_lg_fast(const ProjK &k, const Tab &tab, const ProjPatient &p, double den)
{
    // log(K/den + eps) for den > 0
    return fma(den, p.eps_over_K, fm::base_minus_log_entry(k.lt, tab(fm::log_tab_index(den)), p.logK, den));
}
__device__ __forceinline__ double proj_finish(const ProjPatient &p, double V, double lg, double C, double noise)
{
    double s = fma(p.rho, lg, 1.0);
    s = fma(p.neg_beta_c, C, s);
    s = __dadd_rn(s, noise);
    return __dmul_rn(V, s);
}
// td = alpha d + beta d^2 of the step with radiotherapy
__device__ __forceinline__ double proj_finish_td(const ProjPatient &p, double V, double lg, double C, double td, double noise)
{
    double s = fma(p.rho, lg, 1.0);
    s = fma(p.neg_beta_c, C, s);
    s = __dsub_rn(s, td);
    s = __dadd_rn(s, noise);
    return __dmul_rn(V, s);
}
template <typename Tab>
__device__ __forceinline__ double proj_one(const ProjK &k, const Tab &tab, const ProjPatient &p, double V,
                                           double C, double noise)
{
    const double den = __dadd_rn(V, k.eps);
    double lg = proj_lg_fast(k, tab, p, den);
    if (!(den > 0.0 && den < k.big)) lg = proj_log_slow(p.K, V);     // negative / non-finite volume: reference expression
    return proj_finish(p, V, lg, C, noise);
}
// the chemo chain and the radio chain of an item advance together (two independent dependency chains)
template <typename Tab>
__device__ __forceinline__ void proj_two(const ProjK &k, const Tab &tab, const ProjPatient &p, double &Vc,
                                         double Cc, double &Vr, double Cr, double td, bool radio_now, double noise)
{
    const double dc = __dadd_rn(Vc, k.eps), dr = __dadd_rn(Vr, k.eps);
    double lgc = proj_lg_fast(k, tab, p, dc), lgr = proj_lg_fast(k, tab, p, dr);
    const bool okc = dc > 0.0 && dc < k.big, okr = dr > 0.0 && dr < k.big;
    if (!(okc && okr)) {             // (log of a negative number -> NaN -> the option is dropped, :745-746)
        if (!okc) lgc = proj_log_slow(p.K, Vc);
        if (!okr) lgr = proj_log_slow(p.K, Vr);
    }
    Vc = proj_finish(p, Vc, lgc, Cc, noise);
    Vr = radio_now ? proj_finish_td(p, Vr, lgr, Cr, td, noise) : proj_finish(p, Vr, lgr, Cr, noise);
}

// One work item: the two sliding options that start at projected step s (chemo at s: option s; radio at s: option
// H + s), preceded by s untreated steps.  V0 = F[t+1], C0 = C[t], nz[k] = noise of projected step k.  Writes the
// options' rows rc[0..H), rr[0..H) and returns the validity bits (1: chemo option, 2: radio option).
template <typename Tab, typename Store>
__device__ __forceinline__ unsigned project_item(const ProjK &k, const Tab &tab, const ProjPatient &p, int s,
                                                 double V0, double C0, const double *nz, Store store)
{
    double Vb = V0, Cb = C0;
#pragma unroll 1
    for (int kk = 0; kk < s; ++kk) {                     // untreated steps before the options start
        Cb = __dmul_rn(Cb, k.decay);
        Vb = proj_one(k, tab, p, Vb, Cb, nz[kk]);
        store(kk, Vb, Vb);
    }
    double Cc = __dadd_rn(__dmul_rn(Cb, k.decay), k.chemo_amt);       // chemo dose at step s
    Cb = __dmul_rn(Cb, k.decay);
    double Vc = Vb, Vr = Vb;
    proj_two(k, tab, p, Vc, Cc, Vr, Cb, p.td, true, nz[s]);          // radio (d = radio_amt) at step s
    store(s, Vc, Vr);
#pragma unroll 1
    for (int kk = s + 1; kk < PB_H; ++kk) {
        Cc = __dmul_rn(Cc, k.decay);
        Cb = __dmul_rn(Cb, k.decay);
        proj_two(k, tab, p, Vc, Cc, Vr, Cb, 0.0, false, nz[kk]);
        store(kk, Vc, Vr);
    }
    // a NaN anywhere in an option's volumes drops the option (:745-746); NaN propagates to the last value
    return (isnan(Vc) ? 0u : 1u) | (isnan(Vr) ? 0u : 2u);
}

// ---- phase A ------------------------------------------------------------------------------------------------------
// Patient 0 of the whole cohort reads the row it is writing itself (output row 0 = its first emitted row at t = 0:
// [F0, F1, the H projected volumes of its first valid option, 0, ...]; before that row exists it reads zeros).
// cf_seq_self_kernel runs its t = 0 step and projections ahead of phase A with the very functions phase A and
// phase B use (same operations, same bits) and leaves the H projected volumes of that first option in self_row.
__global__ void cf_seq_self_kernel(int64_t n, int T, int H, SimC2 c, const __grid_constant__ ProjK pk,
                                   const double *__restrict__ params, const double *__restrict__ noise,
                                   const double *__restrict__ chemo_rvs, const double *__restrict__ radio_rvs,
                                   double *__restrict__ self_row, int *__restrict__ err)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const Patient p = load_patient(params, n, 0);
    CfFactual s;
    s.F = p.v0; s.Cprev = 0.0; s.cnt = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) s.win[q] = 0.0;
    double C_t, D_t;
    cf_assign(c, p, s, 0.0, chemo_rvs[0], radio_rvs[0], 0, C_t, D_t);     // the window row is still all zeros
    const double lg = cf_log_ratio(p.K, s.F);
    const double Fn = clip(growth(p, s.F, lg, C_t, D_t, noise[1]), 0.0, c.death);
    const ProjPatient pp = proj_patient(pk, c.radio_amt, params, n, 0);
    const ComputedLogTab tab;
    unsigned vmask = 0;
    double opt[2 * PB_H][PB_H];
    for (int sft = 0; sft < PB_H; ++sft) {
        const unsigned vb = project_item(pk, tab, pp, sft, Fn, C_t, noise + 2,
                                         [&](int kk, double a, double r) { opt[sft][kk] = a; opt[PB_H + sft][kk] = r; });
        vmask |= ((vb & 1u) << sft) | (((vb >> 1) & 1u) << (PB_H + sft));
    }
    if (vmask == 0) { atomicExch(err, 2); return; }   // row 0 would be written by a later step: not modelled
    const int o = __ffs(vmask) - 1;
    for (int kk = 0; kk < PB_H; ++kk) self_row[kk] = opt[o][kk];
}

__global__ void __launch_bounds__(128, CF_FACTUAL_MINB)
cf_seq_factual_kernel(int64_t lo, int64_t hi, int64_t n, int T, int H, SimC2 c, const double *__restrict__ params,
                      const double *__restrict__ noise, const double *__restrict__ rec,
                      const double *__restrict__ chemo_rvs, const double *__restrict__ radio_rvs, int64_t base,
                      CfSrc src, int src_required, const double *__restrict__ self_row, double *__restrict__ F_out,
                      uint8_t *__restrict__ codes_out, double *__restrict__ C_out, int *__restrict__ n_steps,
                      int *__restrict__ err)
{
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const int64_t gi = base + i;
    const int NW = T + H;
    const Patient p = load_patient(params, n, i);
    double *Fr = F_out + i * T;
    uint8_t *cr = codes_out + i * T;
    double *Cr = C_out + i * T;
    WindowRow w;
    w.F = nullptr; w.n_f = 0; w.tail_len = 0; w.self = (gi == 0);
#pragma unroll
    for (int q = 0; q < MAXH; ++q) w.tail[q] = 0.0;
    bool missing = false;
    if (!w.self) {
        const int64_t j = find_owner(src.off, src.n, gi);
        if (j < 0) {
            if (src_required) { atomicExch(err, 1); missing = true; }
            // otherwise the row has not been written yet: the reference reads zeros
        } else {
            int r = (int)(gi - src.off[j]);
            int ts = 0, o = 0;
            for (ts = 0; ts < T - 1; ++ts) {     // locate (step, option) of the owner's r-th emitted row
                const unsigned m = src.valid[j * (T - 1) + ts];
                const int cnt = __popc(m);
                if (r < cnt) {
                    unsigned mm = m;
                    for (int q = 0; q < r; ++q) mm &= mm - 1;
                    o = __ffs(mm) - 1;
                    break;
                }
                r -= cnt;
            }
            w.F = src.F + j * T;
            w.n_f = ts + 2;                      // F[:ts+2] ++ the H projected volumes
            w.tail_len = H;
            for (int q = 0; q < H && q < MAXH; ++q) w.tail[q] = src.cf[((j * (T - 1) + ts) * (2 * H) + o) * H + q];
        }
    }
    CfFactual s;
    s.F = p.v0; s.Cprev = 0.0; s.cnt = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) s.win[q] = 0.0;
    Fr[0] = p.v0;
    double self_F1 = 0.0, self_tail[PB_H];
#pragma unroll
    for (int q = 0; q < PB_H; ++q) self_tail[q] = 0.0;
    int steps = 0;
    // draws of step t + 1 requested while step t computes (see cf_one_step_kernel)
    double nx_chemo = chemo_rvs[i * T], nx_radio = radio_rvs[i * T], nx_rec = rec[i * T], nx_noise = noise[i * NW + 1];
    double nx_w = (w.self || missing) ? 0.0 : w.at(0);
    for (int t = 0; t < T - 1 && !missing; ++t) {
        const double u_chemo = nx_chemo, u_radio = nx_radio, u_rec = nx_rec, nz = nx_noise, w_pre = nx_w;
        if (t + 1 < T - 1) {
            nx_chemo = chemo_rvs[i * T + t + 1]; nx_radio = radio_rvs[i * T + t + 1];
            nx_rec = rec[i * T + t + 1]; nx_noise = noise[i * NW + t + 2];
            if (!w.self) nx_w = w.at(t + 1);
        }
        double w_t;
        if (w.self) {
            if (t == 1) { s.cnt = 0; window_push(s.win, s.cnt, c.window + 1, cf_diameter(p.v0, c.sphere)); }
            w_t = (t == 1) ? self_F1 : 0.0;
#pragma unroll
            for (int q = 0; q < PB_H; ++q)
                if (t == q + 2) w_t = self_tail[q];
        } else {
            w_t = w_pre;
        }
        double C_t, D_t;
        const int fo = cf_assign(c, p, s, w_t, u_chemo, u_radio, t, C_t, D_t);
        const double lg = cf_log_ratio(p.K, s.F);
        const double Fn = clip(growth(p, s.F, lg, C_t, D_t, nz), 0.0, c.death);
        Fr[t + 1] = Fn;
        cr[t] = (uint8_t)fo;
        Cr[t] = C_t;
        if (w.self && t == 0) {
            // row 0 = the first option this patient emits at t = 0 (cf_seq_self_kernel projected it)
            self_F1 = Fn;
#pragma unroll
            for (int kk = 0; kk < PB_H; ++kk) self_tail[kk] = self_row[kk];
        }
        steps = t + 1;
        s.F = Fn;
        s.Cprev = C_t;
        if (Fn >= c.death || recovery_test<false>(u_rec, Fn, c.density)) break;   // death / recovery ends the trajectory
    }
    for (int t = steps; t < T - 1; ++t) { Fr[t + 1] = 0.0; cr[t] = 0; Cr[t] = 0.0; }   // after the last executed step: zeros
    cr[T - 1] = 0;
    Cr[T - 1] = 0.0;
    n_steps[i] = steps;
}

// Pipeline: per chunk c (four factual steps of the CTA's 32 patients), buffers b = c & 1:
//   load/store warp (lane = patient): inputs of chunk c+2 -> in[b] once every compute warp has finished chunk c
//       (in_full[b]); validity masks and row counts of chunk c; one bulk copy per patient stage[b] -> global;
//       out_empty[b] once the copy engine has read the stage
//   compute warps: wait in_full[b] and out_empty[b], run their two items, out_full[b]
// so no warp ever waits for the whole CTA, and global-memory latency never meets the arithmetic.
template <int MINB>
__global__ void __launch_bounds__(PB_THREADS, MINB)
cf_seq_project_kernel(int64_t lo, int64_t hi, int64_t n, int T, const __grid_constant__ ProjK k, double radio_amt,
                      const double *__restrict__ params, const double *__restrict__ noise,
                      const double *__restrict__ F, const double *__restrict__ Cd, const int *__restrict__ n_steps,
                      double *__restrict__ cf_out, uint16_t *__restrict__ valid_out, int *__restrict__ n_rows)
{
    extern __shared__ __align__(128) unsigned char pb_smem[];
    fm::LogTabEntry *tab_s = reinterpret_cast<fm::LogTabEntry *>(pb_smem + PB_OFF_TAB);
    const SmemLogTab tab{tab_s};
    uint64_t *bars = reinterpret_cast<uint64_t *>(pb_smem + PB_OFF_BARS);
    uint64_t *in_full = bars, *out_full = bars + 2, *out_empty = bars + 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < fm::LOGT_SIZE) tab_s[threadIdx.x] = fm::log_table_entry(threadIdx.x);
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&in_full[b], 32);
            mbar_init(&out_full[b], PB_CWARPS);
            mbar_init(&out_empty[b], 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t i_raw = lo + (int64_t)blockIdx.x * 32 + lane;
    const bool exists = i_raw < hi;
    const int64_t i = exists ? i_raw : hi - 1;
    const int NW = T + PB_H;
    const int nchunks = (T - 1 + PB_TT - 1) / PB_TT;

    if (warp == PB_CWARPS) {
        // ------------------------------------------------ load / store warp ------------------------------------
        const double *Fr = F + i * T, *Cr = Cd + i * T, *nr = noise + i * NW;
        auto load_inputs = [&](int c) {
            const int t0 = c * PB_TT;
            double *dst = reinterpret_cast<double *>(pb_smem + PB_OFF_IN + (c & 1) * PB_IN_BYTES) + lane * PB_IN_STRIDE;
            double v[PB_IN_DOUBLES];
#pragma unroll
            for (int q = 0; q < PB_TT; ++q) {
                v[q] = (t0 + 1 + q < T) ? Fr[t0 + 1 + q] : 0.0;
                v[PB_TT + q] = (t0 + q < T) ? Cr[t0 + q] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < PB_TT + PB_H - 1; ++q) v[2 * PB_TT + q] = (t0 + 2 + q < NW) ? nr[t0 + 2 + q] : 0.0;
#pragma unroll
            for (int q = 0; q < PB_IN_DOUBLES; ++q) dst[q] = v[q];
            mbar_arrive(&in_full[c & 1]);
        };
        load_inputs(0);
        if (nchunks > 1) load_inputs(1);
        int rows = 0;
        for (int c = 0; c < nchunks; ++c) {
            const int b = c & 1, t0 = c * PB_TT;
            mbar_wait(&out_full[b], (c >> 1) & 1);
            const int nt = (T - 1 - t0 < PB_TT) ? (T - 1 - t0) : PB_TT;
            if (exists) {
                const uint8_t *vb = pb_smem + PB_OFF_VALID + b * PB_VALID_BYTES + lane * (PB_TT * PB_H);
                for (int q = 0; q < nt; ++q) {
                    unsigned m = 0;
#pragma unroll
                    for (int s = 0; s < PB_H; ++s) {
                        const unsigned v = vb[q * PB_H + s];
                        m |= ((v & 1u) << s) | (((v >> 1) & 1u) << (PB_H + s));
                    }
                    valid_out[i * (T - 1) + t0 + q] = (uint16_t)m;
                    rows += __popc(m);
                }
                bulk_store_s2g(cf_out + (i * (T - 1) + t0) * PB_BLK,
                               reinterpret_cast<double *>(pb_smem + b * PB_STAGE_BYTES) + lane * PB_PSTRIDE,
                               (uint32_t)(nt * PB_BLK * 8));
            }
            tma_store_commit();
            if (c + 2 < nchunks) load_inputs(c + 2);     // in[b] is free: every compute warp has finished chunk c
            tma_store_wait_read();
            mbar_arrive(&out_empty[b]);
        }
        if (exists) n_rows[i] = rows;
        return;
    }
    // ---------------------------------------------------- compute warps ----------------------------------------
    const ProjPatient p = proj_patient(k, radio_amt, params, n, i);
    const int nst = exists ? n_steps[i] : 0;
    // items of this warp inside a chunk: (t_local, s) twice, 16 projected steps in total
    int tl0, s0, tl1, s1;
    if (warp < 4) { tl0 = warp; s0 = 0; tl1 = warp; s1 = 4; }
    else if (warp < 8) { tl0 = warp - 4; s0 = 1; tl1 = warp - 4; s1 = 3; }
    else { tl0 = 2 * (warp - 8); s0 = 2; tl1 = tl0 + 1; s1 = 2; }
    for (int c = 0; c < nchunks; ++c) {
        const int b = c & 1, t0 = c * PB_TT;
        const unsigned par = (c >> 1) & 1;
        mbar_wait(&in_full[b], par);
        mbar_wait(&out_empty[b], par ^ 1);
        const double *in = reinterpret_cast<const double *>(pb_smem + PB_OFF_IN + b * PB_IN_BYTES) + lane * PB_IN_STRIDE;
        double *my_stage = reinterpret_cast<double *>(pb_smem + b * PB_STAGE_BYTES) + lane * PB_PSTRIDE;
        uint8_t *my_valid = pb_smem + PB_OFF_VALID + b * PB_VALID_BYTES + lane * (PB_TT * PB_H);
#pragma unroll 1
        for (int it = 0; it < 2; ++it) {
            const int tl = it ? tl1 : tl0, s = it ? s1 : s0;
            const int t = t0 + tl;
            if (t >= T - 1) continue;
            double *rc = my_stage + tl * PB_BLK + s * PB_H, *rr = rc + PB_H * PB_H;
            unsigned vb = 0;
            if (t < nst) {
                vb = project_item(k, tab, p, s, in[tl], in[PB_TT + tl], in + 2 * PB_TT + tl,
                                  [&](int kk, double a, double r) { rc[kk] = a; rr[kk] = r; });
            } else {
#pragma unroll
                for (int kk = 0; kk < PB_H; ++kk) { rc[kk] = 0.0; rr[kk] = 0.0; }   // after the last executed step: zeros
            }
            my_valid[tl * PB_H + s] = (uint8_t)vb;
        }
        fence_proxy_async_smem();     // this thread's stage writes -> visible to the bulk copy engine
        __syncwarp();
        if (lane == 0) mbar_arrive(&out_full[b]);
    }
}

}  // namespace b200i
