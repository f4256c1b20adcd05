// sim_cf.cu -- K2 / K3: the two counterfactual generators of the cancer simulator in a compact,
// per-patient representation (see include/b200i.h), plus the expanders to the reference's dense rows.
//
//   K2  simulate_counterfactual_1_step          cancer_simulation.py:435-552
//   K3  simulate_counterfactuals_treatment_seq  cancer_simulation.py:635-760 ('sliding_treatment')
//
// One thread per patient.  Per step the factual recursion costs one log + one cbrt + one exp; K2 adds
// four multiply-adds (all options share log(K/F[t])); K3 adds the 2H projected sequences, which share
// their no-treatment prefix: 5 baseline + 15 chemo + 15 radio = 35 log-steps instead of 50 for H = 5.
//
// Cross-row window (cancer_simulation.py:471 / :671): patient i's treatment probabilities are
// computed from output row i, a row emitted by an earlier patient j(i).  Its content is reconstructed
// from j's compact arrays: a prefix of j's factual trajectory, then a short tail (the counterfactual
// value(s) of that row), then zeros.  32 consecutive patients almost always share j, so the reads of
// F_j broadcast within the warp.
#include "fastmath.cuh"
#include "sim_math.cuh"

namespace b200i {

struct SimC2 {
    double death, density, sphere, chemo_amt, radio_amt, decay;
    int window;
};

struct CfSrc {
    int64_t n;  // patients whose row_offsets are known: off[0..n] valid
    const double *F;
    const uint8_t *codes;
    const double *cf;
    const uint16_t *valid;
    const int64_t *off;
};

constexpr int MAXH = 8;
#ifndef CF_FACTUAL_MINB
#define CF_FACTUAL_MINB 3     // resident CTAs per SM the thread-per-patient factual kernels (K2, K3 phase A) are compiled for
#endif

// The factual step's three transcendental functions and its divisions: the lean versions of csrc/fastmath.cuh (the
// ones K1 runs; <= 1 ulp, same class as libdevice's and numpy's own) wherever their domain allows -- positive, finite,
// normal arguments, |sigmoid argument| <= 700 -- and the library functions otherwise (zero or negative window
// entries, an eradicated tumour): 1300 -> ~500 warp instructions per patient step.
__device__ __forceinline__ double cf_diameter(double v, double sphere)
{
    if (v >= 1e-30 && v <= 1e30) return __dmul_rn(fm::cbrt_fast(fm::div_fast(v, sphere)), 2.0);
    if (v == 0.0) return 0.0;      // the part of the window row that was never written: ((0 / c) ** (1/3)) * 2 = 0
    return calc_diameter(v, sphere);
}
__device__ __forceinline__ double cf_sigmoid(double beta, double metric, double intercept)
{
    // 1.0 / (1.0 + np.exp(-beta * (metric - intercept)))   cancer_simulation.py:322-323
    const double z = __dmul_rn(-beta, __dsub_rn(metric, intercept));
    if (fabs(z) <= 700.0) return fm::rcp_fast(__dadd_rn(1.0, fm::exp_fast(z)));
    return __ddiv_rn(1.0, __dadd_rn(1.0, exp(z)));      // also the NaN route (negative volume in the window)
}
__device__ __forceinline__ double cf_log_ratio(double K, double F)
{
    if (F >= 1e-300 && F <= 1e300 && K >= 1e-300 && K <= 1e300) return fm::log_ratio(K, F);
    return log(__ddiv_rn(K, F));
}

// owner j of global row g: off[j] <= g < off[j+1]; -1 if g is beyond the known rows
__device__ __forceinline__ int64_t find_owner(const int64_t *__restrict__ off, int64_t n, int64_t g)
{
    if (n <= 0 || g >= off[n]) return -1;
    int64_t lo = 0, hi = n;  // off[lo] <= g < off[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (off[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// what output row `gi` looked like when patient gi read it
struct WindowRow {
    const double *F;   // owner's factual trajectory
    int n_f;           // leading entries taken from F
    int tail_len;      // then tail[0..tail_len), then zeros
    double tail[MAXH];
    bool self;         // patient 0: reads the row it writes itself at t = 0
    __device__ __forceinline__ double at(int k) const
    {
        if (k < n_f) return F[k];
        const int m = k - n_f;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < MAXH; ++q)
            if (q == m && q < tail_len) v = tail[q];
        return v;
    }
};

struct CfFactual {
    double F, Cprev;         // F[t], chemo dosage of t-1
    double win[16];
    int cnt;
};

// one factual step shared by K2 and K3: window -> probabilities -> assignment -> dosage.
// Returns the factual option index 2*chemo + radio; C_t, D_t by reference.
__device__ __forceinline__ int cf_assign(const SimC2 &c, const Patient &p, CfFactual &s, double w_t, double uchemo,
                                         double uradio, int t, double &C_t, double &D_t)
{
    window_push(s.win, s.cnt, c.window + 1, cf_diameter(w_t, c.sphere));
    const double metric = np_mean(s.win, s.cnt);
    const double pr = cf_sigmoid(p.radio_beta, metric, p.radio_int);
    const double pc = p.same_sigmoid ? pr : cf_sigmoid(p.chemo_beta, metric, p.chemo_int);
    const bool ra = uradio < pr;   // NaN probability (negative volume in the window) -> no treatment
    const bool ca = uchemo < pc;
    D_t = ra ? c.radio_amt : 0.0;
    const double prev = (t == 0) ? 0.0 : s.Cprev;
    C_t = __dadd_rn(__dmul_rn(prev, c.decay), ca ? c.chemo_amt : 0.0);
    return (ca ? 2 : 0) + (ra ? 1 : 0);
}

// V * (1 + rho*lg - beta_c*C - (alpha*d + beta*d^2) + noise) with lg supplied
__device__ __forceinline__ double growth(const Patient &p, double V, double lg, double C, double D, double noise)
{
    double s = __dadd_rn(1.0, __dmul_rn(p.rho, lg));
    s = __dsub_rn(s, __dmul_rn(p.beta_c, C));
    s = __dsub_rn(s, __dadd_rn(__dmul_rn(p.alpha, D), __dmul_rn(p.beta, __dmul_rn(D, D))));
    s = __dadd_rn(s, noise);
    return __dmul_rn(V, s);
}

__device__ __forceinline__ double clip(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

// ------------------------------------------------------------------------------------------------
// K2
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, CF_FACTUAL_MINB)
cf_one_step_kernel(int64_t lo, int64_t hi, int64_t n, int T, SimC2 c, const double *__restrict__ params,
                   const double *__restrict__ noise, const double *__restrict__ rec,
                   const double *__restrict__ chemo_rvs, const double *__restrict__ radio_rvs, int64_t base,
                   CfSrc src, int src_required, double *__restrict__ F_out, uint8_t *__restrict__ codes_out,
                   double *__restrict__ cf_out, int *__restrict__ n_steps, int *__restrict__ n_rows,
                   int *__restrict__ err)
{
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const int64_t gi = base + i;
    const Patient p = load_patient(params, n, i);
    WindowRow w;
    w.F = nullptr; w.n_f = 0; w.tail_len = 0; w.self = (gi == 0);
#pragma unroll
    for (int q = 0; q < MAXH; ++q) w.tail[q] = 0.0;
    if (!w.self) {
        const int64_t j = find_owner(src.off, src.n, gi);
        if (j < 0) {
            if (src_required) { atomicExch(err, 1); return; }
            // row not written yet: the reference reads zeros
        } else {
            const int r = (int)(gi - src.off[j]);
            const int ts = r >> 2, q = r & 3;
            w.F = src.F + j * T;
            if (q == 0) {
                w.n_f = ts + 2;           // factual snapshot taken at step ts: F[:ts+2]
            } else {
                const int fo = src.codes[j * T + ts];
                const int o = (q - 1) + ((q - 1) >= fo ? 1 : 0);   // q-th non-factual option
                w.n_f = ts + 1;           // F[:ts+1] ++ [counterfactual volume]
                w.tail_len = 1;
                w.tail[0] = src.cf[(j * (T - 1) + ts) * 4 + o];
            }
        }
    }
    CfFactual s;
    s.F = p.v0; s.Cprev = 0.0; s.cnt = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) s.win[q] = 0.0;
    double *Fr = F_out + i * T;
    uint8_t *cr = codes_out + i * T;
    Fr[0] = p.v0;
    double self_F1 = 0.0;
    int steps = 0;
    // the four draws of step t + 1 are requested while step t computes (their addresses do not depend on the
    // trajectory): the first version loaded them where they were used and spent 6.4 of 10 issue slots in
    // long_scoreboard stalls (profiles/r2_k2_one_step_ncu.txt)
    double nx_chemo = chemo_rvs[i * T], nx_radio = radio_rvs[i * T], nx_rec = rec[i * T], nx_noise = noise[i * T + 1];
    double nx_w = w.self ? 0.0 : w.at(0);      // the window row's entry of the next step (owner's row: a remote read)
    for (int t = 0; t < T - 1; ++t) {
        const double u_chemo = nx_chemo, u_radio = nx_radio, u_rec = nx_rec, nz = nx_noise, w_pre = nx_w;
        if (t + 1 < T - 1) {
            nx_chemo = chemo_rvs[i * T + t + 1]; nx_radio = radio_rvs[i * T + t + 1];
            nx_rec = rec[i * T + t + 1]; nx_noise = noise[i * T + t + 2];
            if (!w.self) nx_w = w.at(t + 1);
        }
        double w_t;
        if (w.self) {
            // row 0 is this patient's own t=0 snapshot [F0, F1, 0, ...]; before it exists the row is 0
            if (t == 1) { s.cnt = 0; window_push(s.win, s.cnt, c.window + 1, cf_diameter(p.v0, c.sphere)); }
            w_t = (t == 1) ? self_F1 : 0.0;
        } else {
            w_t = w_pre;
        }
        double C_t, D_t;
        const int fo = cf_assign(c, p, s, w_t, u_chemo, u_radio, t, C_t, D_t);
        const double lg = cf_log_ratio(p.K, s.F);
        const double prevC = (t == 0) ? 0.0 : s.Cprev;
        double Vf = 0.0;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const double Co = __dadd_rn(__dmul_rn(prevC, c.decay), (o & 2) ? c.chemo_amt : 0.0);
            const double Do = (o & 1) ? c.radio_amt : 0.0;
            const double V = growth(p, s.F, lg, Co, Do, nz);
            cf_out[(i * (T - 1) + t) * 4 + o] = V;
            if (o == fo) Vf = V;
        }
        const double Fn = clip(Vf, 0.0, c.death);
        Fr[t + 1] = Fn;
        cr[t] = (uint8_t)fo;
        if (t == 0) self_F1 = Fn;
        steps = t + 1;
        s.F = Fn;
        s.Cprev = C_t;
        if (Fn >= c.death || recovery_test<false>(u_rec, Fn, c.density)) break;   // death / recovery ends the trajectory
    }
    for (int t = steps; t < T - 1; ++t) {   // steps after the last executed one: zeros (the compact arrays are fully defined)
        Fr[t + 1] = 0.0; cr[t] = 0;
        double2 *z = reinterpret_cast<double2 *>(cf_out + (i * (T - 1) + t) * 4);
        z[0] = make_double2(0.0, 0.0); z[1] = make_double2(0.0, 0.0);
    }
    cr[T - 1] = 0;
    n_steps[i] = steps;
    n_rows[i] = 4 * steps;
}

// ------------------------------------------------------------------------------------------------
// K3
// ------------------------------------------------------------------------------------------------
// one projected step, cancer_simulation.py:739-743 (note the two 1e-07 terms and no clipping)
// The 35 projected steps per factual step dominate K3, so the division and the logarithm are the lean ones of
// fastmath.cuh (correctly rounded division, log <= 0.9 ulp) whenever the operands are positive, finite and normal;
// anything else (negative projected volumes give the NaN that invalidates an option, :745-746) takes the library
// path, so the validity masks follow the reference exactly.
__device__ __noinline__ double proj_log_slow(double K, double V)
{
    return log(__dadd_rn(__ddiv_rn(K, __dadd_rn(V, 1e-07)), 1e-07));
}
__device__ __forceinline__ double proj_step(const fm::FmK &fk, const Patient &p, double V, double C, double D, double noise)
{
    const double den = __dadd_rn(V, 1e-07);
    double lg;
    if (den > 1e-200 && den < 1e200) {
        const double x = __dadd_rn(fm::div_fast(p.K, den), 1e-07);
        lg = fm::log_ratio(fk, fm::log_num(x), 1.0);
    } else {
        lg = proj_log_slow(p.K, V);
    }
    return growth(p, V, lg, C, D, noise);
}

constexpr int K3_MINB_DEFAULT = 3;   // measured at 1M patients: 2 -> 23.8 ms, 3 -> 20.1 ms, 4 (spills) -> 23.4 ms

template <int H, int MINB>
__global__ void __launch_bounds__(128, MINB)
cf_treatment_seq_kernel(int64_t lo, int64_t hi, int64_t n, int T, SimC2 c, const double *__restrict__ params,
                        const double *__restrict__ noise, const double *__restrict__ rec,
                        const double *__restrict__ chemo_rvs, const double *__restrict__ radio_rvs, int64_t base,
                        CfSrc src, int src_required, double *__restrict__ F_out, uint8_t *__restrict__ codes_out,
                        double *__restrict__ cf_out, uint16_t *__restrict__ valid_out, int *__restrict__ n_steps,
                        int *__restrict__ n_rows, int *__restrict__ err)
{
    static_assert(H >= 1 && H <= MAXH, "projection horizon");
    // The 2H x H projected volumes of a step (400 B per patient at H = 5) are staged in shared memory and written
    // by the whole warp, one patient's block per instruction (25 lanes x 16 B contiguous): per-thread 8-byte stores
    // 23.6 KB apart kept this kernel at 0.8 TB/s of scattered sector writes.
    extern __shared__ __align__(16) double cf_stage[];           // [blockDim][PITCH]
    constexpr int BLK = 2 * H * H, PITCH = BLK + 1;
    const int lane = threadIdx.x & 31;
    double *my_stage = cf_stage + (size_t)threadIdx.x * PITCH;
    const double *warp_stage = cf_stage + (size_t)(threadIdx.x - lane) * PITCH;
    const int64_t i_raw = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool exists = i_raw < hi;
    const int64_t i = exists ? i_raw : lo;                       // lanes past the end idle through the warp copies
    const int64_t i_warp0 = i_raw - lane;
    const int64_t gi = base + i;
    const int NW = T + H;  // noise row width
    const Patient p = load_patient(params, n, i);
    const fm::FmK fk = fm::consts();
    bool missing = false;
    WindowRow w;
    w.F = nullptr; w.n_f = 0; w.tail_len = 0; w.self = (gi == 0);
#pragma unroll
    for (int q = 0; q < MAXH; ++q) w.tail[q] = 0.0;
    if (!w.self) {
        const int64_t j = find_owner(src.off, src.n, gi);
        if (j < 0) {
            if (src_required && exists) { atomicExch(err, 1); }
            missing = src_required != 0;
        } else {
            int r = (int)(gi - src.off[j]);
            int ts = 0, o = 0;
            for (ts = 0; ts < T - 1; ++ts) {     // locate (step, option) of the patient's r-th emitted row
                const unsigned m = src.valid[j * (T - 1) + ts];
                const int cnt = __popc(m);
                if (r < cnt) {
                    unsigned mm = m;
                    for (int q = 0; q < r; ++q) mm &= mm - 1;   // drop the r lowest set bits
                    o = __ffs(mm) - 1;
                    break;
                }
                r -= cnt;
            }
            w.F = src.F + j * T;
            w.n_f = ts + 2;                      // F[:ts+2] ++ the H projected volumes
            w.tail_len = H;
#pragma unroll
            for (int q = 0; q < H; ++q) w.tail[q] = src.cf[((j * (T - 1) + ts) * (2 * H) + o) * H + q];
        }
    }
    CfFactual s;
    s.F = p.v0; s.Cprev = 0.0; s.cnt = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) s.win[q] = 0.0;
    double *Fr = F_out + i * T;
    uint8_t *cr = codes_out + i * T;
    if (exists) Fr[0] = p.v0;
    double self_F1 = 0.0, self_tail[H];
#pragma unroll
    for (int q = 0; q < H; ++q) self_tail[q] = 0.0;
    int steps = 0, rows = 0;
    bool alive = exists && !missing;
    for (int t = 0; t < T - 1; ++t) {
        const bool live = alive;
        if (!live) {
            if (exists) { Fr[t + 1] = 0.0; cr[t] = 0; valid_out[i * (T - 1) + t] = 0; }
#pragma unroll
            for (int e = 0; e < BLK; ++e) my_stage[e] = 0.0;   // steps after the last executed one: a block of zeros
        } else {
        double w_t;
        if (w.self) {
            // row 0 = first row this patient emits at t = 0: [F0, F1, projections of that option, 0, ...]
            if (t == 1) { s.cnt = 0; window_push(s.win, s.cnt, c.window + 1, cf_diameter(p.v0, c.sphere)); }
            w_t = 0.0;
            if (t == 1) w_t = self_F1;
#pragma unroll
            for (int q = 0; q < H; ++q)
                if (t == q + 2) w_t = self_tail[q];
        } else {
            w_t = w.at(t);
        }
        double C_t, D_t;
        const int fo = cf_assign(c, p, s, w_t, chemo_rvs[i * T + t], radio_rvs[i * T + t], t, C_t, D_t);
        const double lg = cf_log_ratio(p.K, s.F);
        const double Fn = clip(growth(p, s.F, lg, C_t, D_t, noise[i * NW + t + 1]), 0.0, c.death);
        Fr[t + 1] = Fn;
        cr[t] = (uint8_t)fo;

        // ---- 2H sliding-treatment projections (:707-756); noise index of projected step k is t+2+k.
        // Level-major order: at projected step k the untreated baseline, the k+1 chemo options and the k+1 radio
        // options that have started are 2k+3 independent recursions, written back to back so that the scheduler
        // can interleave them (option-major order ran each 5-step chain on its own: 0.1 IPC per warp).
        const double *nz = noise + i * NW + t + 2;
        double Vb[H + 1], Cb[H];     // untreated path and its chemo concentration
        double Vc[H], Cc[H];         // option sft (chemo at step sft): current volume / concentration
        double Vr[H];                // option H + sft (radio at step sft): current volume
        double *cfr = my_stage;
        Vb[0] = Fn;
        unsigned nan_c = 0u, nan_r = 0u;   // options that produced a NaN so far (:745-746)
        unsigned nan_b = 0u;               // bit k: baseline volume Vb[k+1] is NaN
#pragma unroll
        for (int k = 0; k < H; ++k) {
            const double Cprev_b = (k == 0) ? C_t : Cb[k - 1];
            Cb[k] = __dadd_rn(__dmul_rn(Cprev_b, c.decay), 0.0);
            Vb[k + 1] = proj_step(fk, p, Vb[k], Cb[k], 0.0, nz[k]);
            if (isnan(Vb[k + 1])) nan_b |= 1u << k;
#pragma unroll
            for (int sft = 0; sft <= k; ++sft) {
                // chemo option: its own concentration from the step it starts
                const double Vin_c = (sft == k) ? Vb[k] : Vc[sft];
                const double Cin = (sft == k) ? Cprev_b : Cc[sft];
                Cc[sft] = __dadd_rn(__dmul_rn(Cin, c.decay), sft == k ? c.chemo_amt : 0.0);
                Vc[sft] = proj_step(fk, p, Vin_c, Cc[sft], 0.0, nz[k]);
                if (isnan(Vc[sft])) nan_c |= 1u << sft;
                cfr[sft * H + k] = Vc[sft];
                // radio option: baseline concentration, dose at the step it starts
                const double Vin_r = (sft == k) ? Vb[k] : Vr[sft];
                Vr[sft] = proj_step(fk, p, Vin_r, Cb[k], sft == k ? c.radio_amt : 0.0, nz[k]);
                if (isnan(Vr[sft])) nan_r |= 1u << sft;
                cfr[(H + sft) * H + k] = Vr[sft];
            }
            // options that start later still follow the baseline at this step
#pragma unroll
            for (int sft = k + 1; sft < H; ++sft) {
                cfr[sft * H + k] = Vb[k + 1];
                cfr[(H + sft) * H + k] = Vb[k + 1];
            }
        }
        // an option is dropped if any of its H volumes is NaN: its own steps or the baseline steps before its start
        unsigned vmask = 0;
#pragma unroll
        for (int sft = 0; sft < H; ++sft) {
            const unsigned before = nan_b & ((1u << sft) - 1u);
            if (before == 0u && !((nan_c >> sft) & 1u)) vmask |= 1u << sft;
            if (before == 0u && !((nan_r >> sft) & 1u)) vmask |= 1u << (H + sft);
        }
        if (t == 0 && vmask != 0u) {
            // row 0 of patient 0 = its first emitted option (window quirk): keep that option's projections
            const int o = __ffs(vmask) - 1;
#pragma unroll
            for (int k = 0; k < H; ++k) self_tail[k] = cfr[o * H + k];
        }
        valid_out[i * (T - 1) + t] = (uint16_t)vmask;
        rows += __popc(vmask);
        if (t == 0) {
            self_F1 = Fn;
            if (w.self && vmask == 0) atomicExch(err, 2);   // row 0 would be written later: not modelled
        }
        steps = t + 1;
        s.F = Fn;
        s.Cprev = C_t;
        if (Fn >= c.death || recovery_test<false>(rec[i * T + t], Fn, c.density)) alive = false;
        }
        // the warp writes the staged blocks of its live patients: one block per instruction
        const unsigned live_mask = __ballot_sync(0xffffffffu, exists);
        __syncwarp();
        for (unsigned m = live_mask; m != 0u; m &= m - 1u) {
            const int r = __ffs(m) - 1;
            if (lane < BLK / 2) {
                // PITCH is odd: 8-byte loads from the stage, 16-byte store (the global block is 16-byte aligned)
                const double *src_row = warp_stage + (size_t)r * PITCH + 2 * lane;
                double2 v;
                v.x = src_row[0]; v.y = src_row[1];
                double *dst = cf_out + ((i_warp0 + r) * (T - 1) + t) * BLK;
                __stcs(reinterpret_cast<double2 *>(dst) + lane, v);
            }
        }
        __syncwarp();
    }
    if (exists) {
        cr[T - 1] = 0;
        n_steps[i] = steps;
        n_rows[i] = rows;
    }
}

}  // namespace b200i
#include "tma.cuh"
#include "sim_cf_seq.cuh"
namespace b200i {

// ------------------------------------------------------------------------------------------------
// exclusive scan of n_rows[lo..hi) continuing from off[lo]; single CTA
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) scan_rows_kernel(const int *__restrict__ n_rows, int64_t *__restrict__ off,
                                                         int64_t lo, int64_t hi)
{
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = off[lo];
    __syncthreads();
    for (int64_t start = lo; start < hi; start += 1024) {
        const int64_t i = start + tid;
        int64_t v = (i < hi) ? (int64_t)n_rows[i] : 0;
        int64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int64_t wv = s_warp[lane];
            int64_t winc = wv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t up = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += up;
            }
            s_warp[lane] = winc - wv;  // exclusive warp offsets
        }
        __syncthreads();
        const int64_t carry = s_carry;
        const int64_t total_incl = carry + s_warp[warp] + incl;
        if (i < hi) off[i + 1] = total_incl;
        __syncthreads();
        if (tid == 1023) s_carry = total_incl;
        __syncthreads();
    }
}

// the same for long ranges, in three launches: per-CTA sums, scan of the sums (one CTA), per-CTA rescan with its offset
constexpr int SCAN_CHUNK = 4096;   // entries per CTA (1024 threads x 4)
__device__ __forceinline__ int64_t block_incl_scan(int64_t v, int64_t *s_warp, int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int64_t wv = s_warp[lane];
        int64_t winc = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t up = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += up;
        }
        s_warp[lane] = winc - wv;
    }
    __syncthreads();
    const int64_t r = incl + s_warp[warp];
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(1024) scan_chunk_sums_kernel(const int *__restrict__ n_rows, int64_t lo, int64_t hi,
                                                               int64_t *__restrict__ sums)
{
    __shared__ int64_t s_warp[32];
    const int64_t first = lo + (int64_t)blockIdx.x * SCAN_CHUNK + 4 * threadIdx.x;
    int64_t v = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (first + q < hi) v += n_rows[first + q];
    const int64_t incl = block_incl_scan(v, s_warp, threadIdx.x);
    if (threadIdx.x == 1023) sums[blockIdx.x] = incl;
}
__global__ void __launch_bounds__(1024) scan_sums_kernel(int64_t *__restrict__ sums, int nb, const int64_t *__restrict__ off_lo)
{
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = *off_lo;
    __syncthreads();
    for (int start = 0; start < nb; start += 1024) {
        const int b = start + threadIdx.x;
        const int64_t v = b < nb ? sums[b] : 0;
        const int64_t incl = block_incl_scan(v, s_warp, threadIdx.x);
        const int64_t carry = s_carry;
        if (b < nb) sums[b] = carry + incl - v;          // exclusive, continuing from off[lo]
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + incl;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024) scan_chunk_apply_kernel(const int *__restrict__ n_rows, int64_t *__restrict__ off,
                                                                int64_t lo, int64_t hi, const int64_t *__restrict__ sums)
{
    __shared__ int64_t s_warp[32];
    const int64_t first = lo + (int64_t)blockIdx.x * SCAN_CHUNK + 4 * threadIdx.x;
    int64_t r[4], v = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        r[q] = (first + q < hi) ? (int64_t)n_rows[first + q] : 0;
        v += r[q];
    }
    const int64_t incl = block_incl_scan(v, s_warp, threadIdx.x);
    int64_t run = sums[blockIdx.x] + incl - v;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        run += r[q];
        if (first + q < hi) off[first + q + 1] = run;
    }
}
// off[i + 1] for i in [lo, hi), continuing from off[lo]
static int scan_rows(const int *n_rows, int64_t *off, int64_t lo, int64_t hi, cudaStream_t st)
{
    if (hi - lo <= 4 * SCAN_CHUNK) {
        scan_rows_kernel<<<1, 1024, 0, st>>>(n_rows, off, lo, hi);
        return check_cuda(cudaGetLastError(), "scan_rows launch");
    }
    const int nb = (int)((hi - lo + SCAN_CHUNK - 1) / SCAN_CHUNK);
    int64_t *sums = nullptr;
    int rc = pool_alloc(reinterpret_cast<void **>(&sums), (size_t)nb * sizeof(int64_t), st);
    if (rc) return rc;
    scan_chunk_sums_kernel<<<nb, 1024, 0, st>>>(n_rows, lo, hi, sums);
    scan_sums_kernel<<<1, 1024, 0, st>>>(sums, nb, off + lo);
    scan_chunk_apply_kernel<<<nb, 1024, 0, st>>>(n_rows, off, lo, hi, sums);
    rc = check_cuda(cudaGetLastError(), "scan_rows launch");
    pool_free(sums, st);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// expanders: one warp per reference row
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
expand_one_step_kernel(int64_t n, int T, const double *__restrict__ F, const uint8_t *__restrict__ codes,
                       const double *__restrict__ cf, const int64_t *__restrict__ off,
                       const double *__restrict__ ptypes, int64_t row_begin, int64_t row_end,
                       double *__restrict__ vol, double *__restrict__ chemo, double *__restrict__ radio,
                       double *__restrict__ seq_len, double *__restrict__ pt_rows)
{
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = row_begin + wid; g < row_end; g += nw) {
        const int64_t j = find_owner(off, n, g);
        if (j < 0) continue;
        const int r = (int)(g - off[j]);
        const int t = r >> 2, q = r & 3;
        const int fo = codes[j * T + t];
        const int o = (q == 0) ? fo : (q - 1) + ((q - 1) >= fo ? 1 : 0);
        const double tail = (q == 0) ? F[j * T + t + 1] : cf[(j * (T - 1) + t) * 4 + o];
        const int64_t orow = (g - row_begin) * T;
        for (int k = lane; k < T; k += 32) {
            double v = 0.0, ca = 0.0, ra = 0.0;
            if (k <= t) v = F[j * T + k];
            else if (k == t + 1) v = tail;
            if (k < t) { const int cd = codes[j * T + k]; ca = (cd >> 1) & 1; ra = cd & 1; }
            else if (k == t) { ca = (o >> 1) & 1; ra = o & 1; }
            vol[orow + k] = v; chemo[orow + k] = ca; radio[orow + k] = ra;
        }
        if (lane == 0) { seq_len[g - row_begin] = (double)(t + 1); pt_rows[g - row_begin] = ptypes[j]; }
    }
}

__global__ void __launch_bounds__(256)
expand_treatment_seq_kernel(int64_t n, int T, int H, const double *__restrict__ F, const uint8_t *__restrict__ codes,
                            const double *__restrict__ cf, const uint16_t *__restrict__ valid,
                            const int64_t *__restrict__ off, const double *__restrict__ ptypes, int64_t row_begin,
                            int64_t row_end, double *__restrict__ vol, double *__restrict__ chemo,
                            double *__restrict__ radio, double *__restrict__ seq_len, double *__restrict__ pt_rows,
                            double *__restrict__ pids, double *__restrict__ pcur)
{
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int Wd = T + H;
    for (int64_t g = row_begin + wid; g < row_end; g += nw) {
        const int64_t j = find_owner(off, n, g);
        if (j < 0) continue;
        int r = (int)(g - off[j]);
        int t = 0, o = 0;
        for (t = 0; t < T - 1; ++t) {
            const unsigned m = valid[j * (T - 1) + t];
            const int cnt = __popc(m);
            if (r < cnt) {
                unsigned mm = m;
                for (int q = 0; q < r; ++q) mm &= mm - 1;
                o = __ffs(mm) - 1;
                break;
            }
            r -= cnt;
        }
        const double *cfo = cf + ((j * (T - 1) + t) * (2 * H) + o) * H;
        const int64_t orow = (g - row_begin) * Wd;
        for (int k = lane; k < Wd; k += 32) {
            double v = 0.0, ca = 0.0, ra = 0.0;
            if (k <= t + 1) v = F[j * T + k];
            else if (k <= t + 1 + H) v = cfo[k - t - 2];
            if (k <= t) { const int cd = codes[j * T + k]; ca = (cd >> 1) & 1; ra = cd & 1; }
            else if (k <= t + H) {
                const int m = k - t - 1;
                ca = (o < H && m == o) ? 1.0 : 0.0;
                ra = (o >= H && m == o - H) ? 1.0 : 0.0;
            }
            vol[orow + k] = v; chemo[orow + k] = ca; radio[orow + k] = ra;
        }
        if (lane == 0) {
            seq_len[g - row_begin] = (double)(t + H + 1);
            pt_rows[g - row_begin] = ptypes[j];
            pids[g - row_begin] = (double)j;
            pcur[g - row_begin] = (double)t;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host: dependency-level driver
// ------------------------------------------------------------------------------------------------
template <typename Launch>
static int run_levels(int64_t n, int64_t base, const b200i_cf_source *source, int *n_rows, int64_t *off,
                      int64_t *total_rows_host, int *levels_host, cudaStream_t st, Launch launch)
{
    // validate before allocating; every later exit goes through the cudaFreeAsync below
    B200I_REQUIRE(source != nullptr || base == 0, B200I_E_ARG, "sim_cf: a shard (global_base > 0) needs a source cohort");
    int *d_err = nullptr;   // scratch: error flag + 8 doubles for the generator's own use (16-byte aligned)
    {
        int rc0 = pool_alloc(reinterpret_cast<void **>(&d_err), 128, st);
        if (rc0) return rc0;
    }
    int levels = 0;
    int64_t total = 0;
    int rc = check_cuda(cudaMemsetAsync(d_err, 0, sizeof(int), st), "memset err");
    if (!rc) rc = check_cuda(cudaMemsetAsync(off, 0, sizeof(int64_t), st), "memset off");
    if (rc) {
    } else if (source != nullptr) {
        rc = launch(0, n, true, d_err);
        if (!rc) {
            rc = scan_rows(n_rows, off, 0, n, st);
        }
        levels = 1;
    } else {
        int64_t lo = 0, hi = (n > 0) ? 1 : 0;
        while (lo < n && !rc) {
            rc = launch(lo, hi, false, d_err);
            if (rc) break;
            rc = scan_rows(n_rows, off, lo, hi, st);
            if (rc) break;
            int64_t off_hi = 0;
            rc = check_cuda(cudaMemcpyAsync(&off_hi, off + hi, sizeof(int64_t), cudaMemcpyDeviceToHost, st), "copy off");
            if (rc) break;
            rc = check_cuda(cudaStreamSynchronize(st), "level sync");
            if (rc) break;
            ++levels;
            lo = hi;
            int64_t next = off_hi;          // patients whose row already exists
            if (next <= lo) next = lo + 1;   // (row not written yet: the reference reads zeros)
            hi = next < n ? next : n;
        }
    }
    int h_err = 0;
    if (!rc) rc = check_cuda(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st), "copy err");
    if (!rc && (total_rows_host || true))
        rc = check_cuda(cudaMemcpyAsync(&total, off + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st), "copy total");
    if (!rc) rc = check_cuda(cudaStreamSynchronize(st), "final sync");
    pool_free(d_err, st);
    if (rc) return rc;
    B200I_REQUIRE(h_err != 1, B200I_E_ARG, "sim_cf: source cohort does not cover the rows this shard reads");
    B200I_REQUIRE(h_err != 2, B200I_E_UNSUPPORTED, "sim_cf: patient 0 emitted no row at t=0 (not modelled)");
    B200I_REQUIRE(h_err != 3, B200I_E_UNSUPPORTED, "sim_cf: internal: patient 0 reached the two-phase kernels");
    if (total_rows_host) *total_rows_host = total;
    if (levels_host) *levels_host = levels;
    return 0;
}

static int check_consts(const b200i_sim_consts *k, int T, const char *who)
{
    B200I_REQUIRE(k != nullptr, B200I_E_ARG, "%s: consts is NULL", who);
    B200I_REQUIRE(T >= 3 && T <= 4096, B200I_E_UNSUPPORTED, "%s: seq_length %d outside [3,4096]", who, T);
    B200I_REQUIRE(k->lag == 0, B200I_E_UNSUPPORTED, "%s: lag=%d (only lag=0 is implemented)", who, k->lag);
    B200I_REQUIRE(k->window_size >= 1 && k->window_size <= 15, B200I_E_UNSUPPORTED, "%s: window_size=%d outside [1,15]",
                  who, k->window_size);
    return 0;
}

}  // namespace b200i

using namespace b200i;

extern "C" int b200i_sim_cf_one_step(int64_t n, int32_t T, const b200i_sim_consts *k, const double *params,
                                     const double *noise, const double *recovery_rvs, const double *chemo_rvs,
                                     const double *radio_rvs, int64_t global_base, const b200i_cf_source *source,
                                     double *factual, uint8_t *codes, double *cf, int32_t *n_steps, int32_t *n_rows,
                                     int64_t *row_offsets, int64_t *total_rows_host, int32_t *levels_host, void *stream)
{
    B200I_REQUIRE(n >= 0 && params && noise && recovery_rvs && chemo_rvs && radio_rvs && factual && codes && cf &&
                      n_steps && n_rows && row_offsets,
                  B200I_E_ARG, "sim_cf_one_step: NULL argument or negative n");
    int rc = check_consts(k, T, "sim_cf_one_step");
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SimC2 c{k->death_threshold, k->cell_density, k->sphere_coef, k->chemo_amt, k->radio_amt, k->drug_decay,
            k->window_size};
    auto launch = [&](int64_t lo, int64_t hi, bool required, int *d_err) -> int {
        if (hi <= lo) return 0;
        CfSrc src;
        if (source) src = CfSrc{source->n, source->factual, source->codes, source->cf, nullptr, source->row_offsets};
        else src = CfSrc{lo, factual, codes, cf, nullptr, row_offsets};
        const unsigned grid = (unsigned)((hi - lo + 127) / 128);
        cf_one_step_kernel<<<grid, 128, 0, st>>>(lo, hi, n, T, c, params, noise, recovery_rvs, chemo_rvs, radio_rvs,
                                                 global_base, src, required ? 1 : 0, factual, codes, cf, n_steps,
                                                 n_rows, d_err);
        return check_cuda(cudaGetLastError(), "cf_one_step launch");
    };
    return run_levels(n, global_base, source, n_rows, row_offsets, total_rows_host, levels_host, st, launch);
}

extern "C" int b200i_sim_cf_treatment_seq(int64_t n, int32_t T, int32_t H, const b200i_sim_consts *k,
                                          const double *params, const double *noise, const double *recovery_rvs,
                                          const double *chemo_rvs, const double *radio_rvs, int64_t global_base,
                                          const b200i_cf_source *source, double *factual, uint8_t *codes, double *cf,
                                          uint16_t *valid, int32_t *n_steps, int32_t *n_rows, int64_t *row_offsets,
                                          int64_t *total_rows_host, int32_t *levels_host, void *stream)
{
    B200I_REQUIRE(n >= 0 && params && noise && recovery_rvs && chemo_rvs && radio_rvs && factual && codes && cf &&
                      valid && n_steps && n_rows && row_offsets,
                  B200I_E_ARG, "sim_cf_treatment_seq: NULL argument or negative n");
    int rc = check_consts(k, T, "sim_cf_treatment_seq");
    if (rc) return rc;
    B200I_REQUIRE(H == 5, B200I_E_UNSUPPORTED,
                  "sim_cf_treatment_seq: projection_horizon=%d (the kernel is instantiated for 5, the only value "
                  "the reference configures: config/dataset/cancer_sim.yaml:16)", H);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SimC2 c{k->death_threshold, k->cell_density, k->sphere_coef, k->chemo_amt, k->radio_amt, k->drug_decay,
            k->window_size};
    B200I_REQUIRE(aligned16(cf), B200I_E_ALIGN, "sim_cf_treatment_seq: cf must be 16-byte aligned");
    // B200I_K3_VARIANT=1: the first-generation kernel for every patient (cross-check); default: two-phase kernels,
    // the first-generation kernel only for patient 0 (its window reads the row it is writing itself)
    static const int variant = getenv("B200I_K3_VARIANT") ? atoi(getenv("B200I_K3_VARIANT")) : 0;
    // resident CTAs per SM the first-generation kernel's register budget is sized for; B200I_K3_MINB overrides
    static const int minb = getenv("B200I_K3_MINB") ? atoi(getenv("B200I_K3_MINB")) : K3_MINB_DEFAULT;
    double *conc = nullptr;   // chemo concentration C[t] of the factual trajectories: phase A -> phase B
    if (variant != 1 && n > 0) {
        int rc0 = pool_alloc(reinterpret_cast<void **>(&conc), (size_t)n * T * sizeof(double), st);
        if (rc0) return rc0;
    }
    auto legacy = [&](int64_t lo, int64_t hi, const CfSrc &src, bool required, int *d_err) -> int {
        const unsigned grid = (unsigned)((hi - lo + 127) / 128);
        constexpr int STAGE_BYTES = 128 * (2 * 5 * 5 + 1) * 8;   // [128 threads][2H*H + 1] doubles
        auto go = [&](auto kern) -> int {
            int rc_ = ensure_dyn_smem(reinterpret_cast<const void *>(kern), STAGE_BYTES, 128, nullptr);
            if (rc_) return rc_;
            kern<<<grid, 128, STAGE_BYTES, st>>>(lo, hi, n, T, c, params, noise, recovery_rvs, chemo_rvs, radio_rvs,
                                                 global_base, src, required ? 1 : 0, factual, codes, cf, valid, n_steps,
                                                 n_rows, d_err);
            return check_cuda(cudaGetLastError(), "cf_treatment_seq launch");
        };
        return minb == 2 ? go(cf_treatment_seq_kernel<5, 2>) : (minb == 4 ? go(cf_treatment_seq_kernel<5, 4>)
                                                                           : go(cf_treatment_seq_kernel<5, 3>));
    };
    auto launch = [&](int64_t lo, int64_t hi, bool required, int *d_err) -> int {
        if (hi <= lo) return 0;
        CfSrc src;
        if (source) src = CfSrc{source->n, source->factual, source->codes, source->cf, source->valid, source->row_offsets};
        else src = CfSrc{lo, factual, codes, cf, valid, row_offsets};
        if (variant == 1) return legacy(lo, hi, src, required, d_err);
        const ProjK pk{fm::log_tab_consts(), 1e-07, 1e300, c.decay, c.chemo_amt};
        double *self_row = reinterpret_cast<double *>(d_err) + 1;   // run_levels' scratch: [error flag | H doubles]
        int rc_ = 0;
        if (global_base + lo == 0) {
            cf_seq_self_kernel<<<1, 32, 0, st>>>(n, T, H, c, pk, params, noise, chemo_rvs, radio_rvs, self_row, d_err);
            rc_ = check_cuda(cudaGetLastError(), "cf_seq_self launch");
            if (rc_) return rc_;
        }
        cf_seq_factual_kernel<<<(unsigned)((hi - lo + 127) / 128), 128, 0, st>>>(
            lo, hi, n, T, H, c, params, noise, recovery_rvs, chemo_rvs, radio_rvs, global_base, src, required ? 1 : 0,
            self_row, factual, codes, conc, n_steps, d_err);
        rc_ = check_cuda(cudaGetLastError(), "cf_seq_factual launch");
        if (rc_) return rc_;
        static const int pb_minb = getenv("B200I_K3_PB_MINB") ? atoi(getenv("B200I_K3_PB_MINB")) : 2;
        auto kern = pb_minb == 1 ? cf_seq_project_kernel<1> : cf_seq_project_kernel<2>;
        rc_ = ensure_dyn_smem(reinterpret_cast<const void *>(kern), PB_SMEM_BYTES, PB_THREADS, nullptr);
        if (rc_) return rc_;
        kern<<<(unsigned)((hi - lo + 31) / 32), PB_THREADS, PB_SMEM_BYTES, st>>>(lo, hi, n, T, pk, c.radio_amt, params, noise,
                                                                                factual, conc, n_steps, cf, valid, n_rows);
        return check_cuda(cudaGetLastError(), "cf_seq_project launch");
    };
    const int rc_levels = run_levels(n, global_base, source, n_rows, row_offsets, total_rows_host, levels_host, st, launch);
    pool_free(conc, st);
    return rc_levels;
}

extern "C" int b200i_expand_cf_one_step(int64_t n, int32_t T, const double *factual, const uint8_t *codes,
                                        const double *cf, const int64_t *row_offsets, const double *patient_types,
                                        int64_t row_begin, int64_t row_end, double *cancer_volume,
                                        double *chemo_application, double *radio_application,
                                        double *sequence_lengths, double *patient_types_rows, void *stream)
{
    B200I_REQUIRE(n >= 0 && factual && codes && cf && row_offsets && patient_types && cancer_volume &&
                      chemo_application && radio_application && sequence_lengths && patient_types_rows &&
                      row_end >= row_begin && row_begin >= 0,
                  B200I_E_ARG, "expand_cf_one_step: bad argument");
    if (row_end == row_begin) return 0;
    int64_t grid = (row_end - row_begin + 7) / 8;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (grid > cap) grid = cap;
    expand_one_step_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n, T, factual, codes, cf, row_offsets, patient_types, row_begin, row_end, cancer_volume, chemo_application,
        radio_application, sequence_lengths, patient_types_rows);
    return check_cuda(cudaGetLastError(), "expand_one_step launch");
}

extern "C" int b200i_expand_cf_treatment_seq(int64_t n, int32_t T, int32_t H, const double *factual,
                                             const uint8_t *codes, const double *cf, const uint16_t *valid,
                                             const int64_t *row_offsets, const double *patient_types,
                                             int64_t row_begin, int64_t row_end, double *cancer_volume,
                                             double *chemo_application, double *radio_application,
                                             double *sequence_lengths, double *patient_types_rows,
                                             double *patient_ids, double *patient_current_t, void *stream)
{
    B200I_REQUIRE(n >= 0 && factual && codes && cf && valid && row_offsets && patient_types && cancer_volume &&
                      chemo_application && radio_application && sequence_lengths && patient_types_rows &&
                      patient_ids && patient_current_t && row_end >= row_begin && row_begin >= 0,
                  B200I_E_ARG, "expand_cf_treatment_seq: bad argument");
    if (row_end == row_begin) return 0;
    int64_t grid = (row_end - row_begin + 7) / 8;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (grid > cap) grid = cap;
    expand_treatment_seq_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n, T, H, factual, codes, cf, valid, row_offsets, patient_types, row_begin, row_end, cancer_volume,
        chemo_application, radio_application, sequence_lengths, patient_types_rows, patient_ids, patient_current_t);
    return check_cuda(cudaGetLastError(), "expand_treatment_seq launch");
}
