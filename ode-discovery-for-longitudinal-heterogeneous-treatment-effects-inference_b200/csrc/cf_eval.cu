// cf_eval.cu -- the discovered ODE evaluated on COMPACT counterfactual cohorts (K8 one-step, K9 treatment sequences).
//
// Reference path: TimeVaryingCausalModel.get_normalised_masked_rmse(test_cf_one_step, one_step_counterfactual=True)
// and get_normalised_n_step_rmses(test_cf_treatment_seq) (time_varying_model.py:236-313) over the predictions of
// SINDY.get_predictions / get_autoregressive_predictions (sindy.py:371-431, 433-715, 717-760).  The reference
// materialises one dense row per (patient, t, option) (0.9 TB at 1M patients); every row of one (patient, t) shares
// the factual prefix F[0..t] / F[0..t+1] and -- in INSITE mode -- the fit problem, so here a cohort is evaluated
// straight from the per-patient arrays the generators K2 / K3 write (include/b200i.h, "compact representation"):
//
//   one-step rows of (i,t)  [4 rows: the factual snapshot + the 3 other options; sequence_length = t+1]
//       prediction k (k <= t) = xhat[k+1], rolled open loop from F[0] over the factual codes 0..t-1 and the row's own
//       option at step t; target F[k+1] for k < t, F[t+1] (factual row) or cf[t][o] at k = t.
//   sequence rows of (i,t)  [<= 2H rows, one per valid sliding option; sequence_length = t+H+1]
//       the H scored predictions are xhat[t+2..t+1+H]: factual codes 0..t, then the option's H codes; targets cf[t][o][0..H).
//
// Coefficients: one (4,4) matrix for the whole cohort (population SINDy; terms with |c| <= drop_below dropped,
// pkpd/utils.py:388) or one per (patient, t) (INSITE: the fit only depends on (patient, t), b200i_insite_bfgs_prefix).
//
// Arithmetic: the right-hand side c0 + c1 x + c2 u + c3 x u is affine in x for a patient (u constant), so the
// `substeps` explicit Euler sub-steps of one interval (pkpd/utils.py:68-90) compose to ONE affine map
// x -> al * x + be per treatment code: al = (1 + h B)^s, be = h A (1 + (1 + h B) + ... ), A = c0 + c2 u, B = c1 + c3 u.
// That is the same polynomial the reference evaluates step by step; the rounding differs at the 1e-16 level per step
// (tests: 8 logged RMSEs to 1e-9, dense K6 path to 1e-12).  A rollout is then one FMA per interval, and the kernels
// are bound by reading the cohort (2.4 KB / 24 KB per patient) -- K9 streams each patient's (T-1,2H,H) block with one
// bulk copy (cp.async.bulk, SASS UBLKCP) into a 2-stage shared-memory ring (two stages leave room for four CTAs per SM: 6.0 -> 5.1 ms at 1M patients against three stages and three CTAs).
#include "stats_reduce.cuh"
#include "tma.cuh"

namespace b200i {

constexpr int EV_THREADS = 128;
constexpr int EV_WARPS = EV_THREADS / 32;
constexpr int EV_STAGES = 2;
constexpr int EV_MAXH = 8;
constexpr int EV_MAXT = 128;            // T <= 128
constexpr int EV_SL = EV_MAXT / 32;     // columns per lane
constexpr int EV_MAX_BLOCKS = 2048;

struct Affine {
    double al, be;
};

__device__ __forceinline__ double keep_term(double c, double drop_below) { return fabs(c) > drop_below ? c : 0.0; }

// `substeps` explicit Euler sub-steps of x' = (c0 + c2 u) + (c1 + c3 u) x with step h, as one affine map
__device__ __forceinline__ Affine euler_affine(double c0, double c1, double c2, double c3, double u, double h, int substeps)
{
    const double A = c0 + c2 * u, B = c1 + c3 * u;
    const double m = 1.0 + h * B, c = h * A;
    Affine r{1.0, 0.0};
    for (int s = 0; s < substeps; ++s) {
        r.be = fma(m, r.be, c);
        r.al *= m;
    }
    return r;
}
__device__ __forceinline__ double apply(const Affine &m, double v) { return fma(m.al, v, m.be); }
// second after first
__device__ __forceinline__ Affine compose(const Affine &second, const Affine &first)
{
    return Affine{second.al * first.al, fma(second.al, first.be, second.be)};
}
// compact option index 2*chemo + radio (cancer_simulation.py:513) -> coefficient row chemo + 2*radio (dataset.py:127-141)
__device__ __forceinline__ int option_to_code(int o) { return ((o & 1) << 1) | ((o >> 1) & 1); }

__device__ __forceinline__ Affine pick(const Affine (&m)[4], int a)
{
    return a == 0 ? m[0] : (a == 1 ? m[1] : (a == 2 ? m[2] : m[3]));
}

// ordered, atomics-free grid combine of `nvals` per-block values (masked_se's pattern)
__device__ __forceinline__ bool grid_arrive(unsigned int *ticket, unsigned int *s_flag)
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        *s_flag = (t == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (*s_flag) __threadfence();
    return *s_flag != 0u;
}

// ------------------------------------------------------------------------------------------------
// K8: one-step cohort.  One warp per patient.
// sums layout = b200i_masked_se's: se_col[W], cnt_col[W], se_last_col[W], total last se, rows.
// ------------------------------------------------------------------------------------------------
// The inputs of the NEXT patient of a warp are fetched into registers (TW / 32 volumes and code bytes, TW / 8
// counterfactual values per lane) before the current one is evaluated, so the global-load latency overlaps the
// arithmetic; the first version loaded, parked and then computed, and stalled on `long_scoreboard` 8.7 of every 10 issue
// slots (profiles/r2_k8_cf_eval_one_step_ncu.txt).  TW = 64 or 128: compile-time bound of T (register arrays).
template <bool PER_T, int TW>
__global__ void __launch_bounds__(EV_THREADS)
cf_eval_one_step_kernel(int64_t n, int T, double h, int substeps, const double *__restrict__ F,
                        const uint8_t *__restrict__ codes, const double *__restrict__ cf,
                        const int *__restrict__ n_steps, const double *__restrict__ static_u,
                        const double *__restrict__ coefs, double drop_below, double *__restrict__ partials,
                        unsigned int *__restrict__ ticket, double *__restrict__ sums)
{
    // padded layouts: a lane of the shared-coefficient path owns 4 consecutive columns, i.e. 16 consecutive cf values
    // and 4 consecutive F values; pitches 17 / 5 keep the lanes of a half-warp in different banks
    __shared__ double s_F[EV_WARPS][5 * (EV_MAXT / 4) + 8];
    __shared__ double s_cf[EV_WARPS][17 * (EV_MAXT / 4)];
    __shared__ uint8_t s_code[EV_WARPS][EV_MAXT];
    __shared__ Affine s_map[EV_WARPS][4];
    __shared__ double s_coef[16];
    __shared__ double s_acc[EV_WARPS][3 * EV_MAXT];
    auto f_at = [](int k) { return 5 * (k >> 2) + (k & 3); };
    auto cf_at = [](int k, int o) { return 17 * (k >> 2) + 4 * (k & 3) + o; };
    __shared__ unsigned int s_flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = T - 1;
    if (!PER_T && tid < 16) s_coef[tid] = keep_term(coefs[tid], drop_below);
    __syncthreads();

    // PER_T : lane owns columns lane + 32 j        (se[j], cnt[j], last[j])
    // shared: lane owns columns 4 lane + q         (same arrays, index q)
    double se[EV_SL], cnt[EV_SL], last[EV_SL];
#pragma unroll
    for (int j = 0; j < EV_SL; ++j) se[j] = cnt[j] = last[j] = 0.0;

    const int64_t nwarps = (int64_t)gridDim.x * EV_WARPS;
    constexpr int SLF = TW / 32, SLC = TW / 8;
    double rF[SLF], rCF[SLC], r_u = 0.0;
    uint8_t rC[SLF];
    int r_ns = 0;
    auto fetch = [&](int64_t i) {
#pragma unroll
        for (int j = 0; j < SLF; ++j) {
            const int k = lane + 32 * j;
            if (k < T) { rF[j] = F[i * T + k]; rC[j] = codes[i * T + k]; }
        }
#pragma unroll
        for (int j = 0; j < SLC; ++j) {
            const int e = lane + 32 * j;
            if (e < 4 * W) rCF[j] = cf[i * (int64_t)(4 * W) + e];
        }
        r_ns = n_steps[i];
        r_u = static_u[i];
    };
    {
        const int64_t i0 = (int64_t)blockIdx.x * EV_WARPS + warp;
        if (i0 < n) fetch(i0);
    }
    for (int64_t i = (int64_t)blockIdx.x * EV_WARPS + warp; i < n; i += nwarps) {
        int ns = r_ns;
        if (ns > W) ns = W;
        const double u = r_u;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < SLF; ++j) {
            const int k = lane + 32 * j;
            if (k < T) { s_F[warp][f_at(k)] = rF[j]; s_code[warp][k] = rC[j]; }
        }
#pragma unroll
        for (int j = 0; j < SLC; ++j) {
            const int e = lane + 32 * j;
            if (e < 4 * W) s_cf[warp][17 * (e >> 4) + (e & 15)] = rCF[j];
        }
        if (i + nwarps < n) fetch(i + nwarps);
        if (!PER_T && lane < 4)   // the four maps of this patient, by option index 2*chemo + radio
            s_map[warp][lane] = euler_affine(s_coef[4 * option_to_code(lane)], s_coef[4 * option_to_code(lane) + 1],
                                             s_coef[4 * option_to_code(lane) + 2], s_coef[4 * option_to_code(lane) + 3], u, h,
                                             substeps);
        __syncwarp();
        const double F0 = s_F[warp][0];
        if (!PER_T) {
            const Affine *mo = s_map[warp];
            // affine prefix scan over the factual steps: lane owns steps 4 lane .. 4 lane + 3
            Affine comp{1.0, 0.0};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = 4 * lane + q;
                if (k < ns) comp = compose(mo[s_code[warp][k] & 3], comp);
            }
            Affine inc = comp;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                Affine prev;
                prev.al = __shfl_up_sync(0xffffffffu, inc.al, d);
                prev.be = __shfl_up_sync(0xffffffffu, inc.be, d);
                if (lane >= d) inc = compose(inc, prev);
            }
            Affine exc;
            exc.al = __shfl_up_sync(0xffffffffu, inc.al, 1);
            exc.be = __shfl_up_sync(0xffffffffu, inc.be, 1);
            if (lane == 0) exc = Affine{1.0, 0.0};
            double v = apply(exc, F0);   // xhat[4 lane]
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = 4 * lane + q;
                if (k < ns) {
                    const int fo = s_code[warp][k] & 3;
                    double e_f = 0.0, e_o = 0.0;
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        const double p = apply(mo[o], v);
                        const double target = (o == fo) ? s_F[warp][f_at(k + 1)] : s_cf[warp][cf_at(k, o)];
                        const double d = p - target;
                        if (o == fo) e_f = d * d;
                        else e_o += d * d;
                    }
                    // rows (i, t > k) all carry the factual step k: 4 (ns-1-k) rows; the 4 rows of t = k end here
                    se[q] += e_f * (double)(4 * (ns - 1 - k) + 1) + e_o;
                    cnt[q] += (double)(4 * (ns - k));
                    last[q] += e_f + e_o;
                    v = apply(mo[fo], v);
                }
            }
        } else {
            const int npass = (ns + 31) >> 5;
            for (int p = 0; p < npass; ++p) {
                const int t = lane + 32 * p;
                const bool act = t < ns;
                Affine m[4];
                if (act) {
                    const double *c = coefs + (i * W + t) * 16;
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                        m[a] = euler_affine(keep_term(c[4 * a], drop_below), keep_term(c[4 * a + 1], drop_below),
                                            keep_term(c[4 * a + 2], drop_below), keep_term(c[4 * a + 3], drop_below), u, h,
                                            substeps);
                } else {
#pragma unroll
                    for (int a = 0; a < 4; ++a) m[a] = Affine{1.0, 0.0};
                }
                const int kmax = (ns < 32 * p + 32) ? ns : 32 * p + 32;
                double v = F0;
                for (int k = 0; k < kmax; ++k) {
                    const int fo = s_code[warp][k] & 3;
                    if (act && k == t) {
                        double e_all = 0.0;
#pragma unroll
                        for (int o = 0; o < 4; ++o) {
                            const double pr = apply(pick(m, option_to_code(o)), v);
                            const double target = (o == fo) ? s_F[warp][f_at(k + 1)] : s_cf[warp][cf_at(k, o)];
                            const double d = pr - target;
                            e_all += d * d;
                        }
#pragma unroll
                        for (int j = 0; j < EV_SL; ++j)
                            if (j == p) { se[j] += e_all; last[j] += e_all; }
                    }
                    v = apply(pick(m, option_to_code(fo)), v);
                    const double d = v - s_F[warp][f_at(k + 1)];
                    const double tot = warp_sum((act && k < t) ? 4.0 * d * d : 0.0);
                    if (lane == (k & 31)) {
#pragma unroll
                        for (int j = 0; j < EV_SL; ++j)
                            if (j == (k >> 5)) se[j] += tot;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < EV_SL; ++j) {
                const int k = lane + 32 * j;
                if (k < ns) cnt[j] += (double)(4 * (ns - k));
            }
        }
    }
    // block combine (columns in natural order), then ordered grid combine
#pragma unroll
    for (int j = 0; j < EV_SL; ++j) {
        const int k = PER_T ? lane + 32 * j : 4 * lane + j;
        s_acc[warp][k] = se[j];
        s_acc[warp][EV_MAXT + k] = cnt[j];
        s_acc[warp][2 * EV_MAXT + k] = last[j];
    }
    __syncthreads();
    constexpr int NV = 3 * EV_MAXT;
    for (int j = tid; j < NV; j += EV_THREADS) {
        double v = 0.0;
        for (int w = 0; w < EV_WARPS; ++w) v += s_acc[w][j];
        partials[(size_t)blockIdx.x * NV + j] = v;
    }
    if (grid_arrive(ticket, &s_flag)) {
        for (int j = tid; j < NV; j += EV_THREADS) {
            double v = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) v += __ldcg(&partials[(size_t)b * NV + j]);
            const int which = j / EV_MAXT, k = j % EV_MAXT;
            if (k < W) sums[which * W + k] = v;
            s_acc[0][j] = v;
        }
        __syncthreads();
        if (tid == 0) {
            double tl = 0.0;
            for (int k = 0; k < W; ++k) tl += s_acc[0][2 * EV_MAXT + k];
            sums[3 * W] = tl;
            sums[3 * W + 1] = s_acc[0][EV_MAXT];   // rows = active count of column 0
            *ticket = 0u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K9: treatment-sequence cohort.  One CTA walks patients; each patient's (T-1, 2H, H) block of projected volumes
// arrives by one bulk copy into a ring of EV_STAGES shared-memory buffers.
// sums layout: se[H] (squared error per projection step), then rows[H] (valid rows, the same for every step).
//
// One __syncthreads per patient: the small per-patient inputs (code bytes, validity masks, executed steps, static
// feature, F[0]) of the NEXT patient are loaded into registers while the current one is processed and parked in a
// double-buffered shared array at the top of the next iteration; that barrier also proves the previous patient's stage
// free, so one thread refills it.  With shared coefficients every warp runs the factual prefix as an affine scan of
// its own (2 steps per lane, 5 shuffle rounds) -- the first version walked 59 dependent FMAs behind two more barriers
// and spent its time in `barrier` / `long_scoreboard` stalls (profiles/r2_k9_cf_eval_seq_ncu.txt: 9.9 ms per 1M patients).
// ------------------------------------------------------------------------------------------------
template <bool PER_T, int HT>   // HT: compile-time projection horizon (0 = run time)
__global__ void __launch_bounds__(EV_THREADS)
cf_eval_seq_kernel(int64_t n, int T, int H_rt, double h, int substeps, const double *__restrict__ F,
                   const uint8_t *__restrict__ codes, const double *__restrict__ cf, const uint16_t *__restrict__ valid,
                   const int *__restrict__ n_steps, const double *__restrict__ static_u, const double *__restrict__ coefs,
                   double drop_below, double *__restrict__ partials, unsigned int *__restrict__ ticket,
                   double *__restrict__ sums)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t full[EV_STAGES];
    __shared__ Affine s_aff[PER_T ? EV_MAXT : 1][4];
    __shared__ double s_start[PER_T ? 1 : EV_WARPS][EV_MAXT];
    __shared__ uint8_t s_code[2][EV_MAXT];
    __shared__ uint16_t s_valid[2][EV_MAXT];
    __shared__ double s_coef[16];
    __shared__ double s_red[EV_WARPS][2 * EV_MAXH];
    __shared__ unsigned int s_flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = HT > 0 ? HT : H_rt;
    const int W = T - 1, O = 2 * H, npairs = W * O;
    const uint32_t block_bytes = (uint32_t)npairs * H * sizeof(double);
    const uint32_t stage_bytes = (block_bytes + 127u) & ~127u;
    const int64_t block_elems = (int64_t)npairs * H;

    if (tid == 0) {
        for (int s = 0; s < EV_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    if (!PER_T && tid < 16) s_coef[tid] = keep_term(coefs[tid], drop_below);
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < EV_STAGES; ++s) {
            const int64_t i = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
            if (i < n) {
                mbar_arrive_expect_tx(&full[s], block_bytes);
                bulk_load_g2s(smem_raw + (size_t)s * stage_bytes, cf + i * block_elems, block_bytes, &full[s]);
            }
        }
    }
    double se[EV_MAXH], cnt = 0.0;
#pragma unroll
    for (int k = 0; k < EV_MAXH; ++k) se[k] = 0.0;

    // small inputs of the first patient
    int r_ns = 0;
    double r_u = 0.0, r_f0 = 0.0;
    uint8_t r_code = 0;
    uint16_t r_valid = 0;
    if ((int64_t)blockIdx.x < n) {
        const int64_t i = blockIdx.x;
        r_ns = n_steps[i]; r_u = static_u[i]; r_f0 = F[i * T];
        if (tid < T) r_code = codes[i * T + tid];
        if (tid < W) r_valid = valid[i * W + tid];
    }
    uint32_t it = 0;
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x, ++it) {
        const int s = (int)(it % EV_STAGES);
        const uint32_t parity = (it / EV_STAGES) & 1u;
        const int b = (int)(it & 1u);
        int ns = r_ns < W ? r_ns : W;
        const double u = r_u, f0 = r_f0;
        if (tid < T) s_code[b][tid] = r_code;
        if (tid < W) s_valid[b][tid] = r_valid;
        {   // next patient's small inputs -> registers (consumed at the top of the next iteration)
            const int64_t nx = i + gridDim.x;
            if (nx < n) {
                r_ns = n_steps[nx]; r_u = static_u[nx]; r_f0 = F[nx * T];
                if (tid < T) r_code = codes[nx * T + tid];
                if (tid < W) r_valid = valid[nx * W + tid];
            }
        }
        __syncthreads();   // small inputs visible; every thread has left the previous patient's stage
        if (tid == 0 && it >= 1) {
            const int64_t nxt = i + (int64_t)(EV_STAGES - 1) * gridDim.x;    // patient it + EV_STAGES - 1
            const int sp = (int)((it - 1) % EV_STAGES);
            if (nxt < n) {
                mbar_arrive_expect_tx(&full[sp], block_bytes);
                bulk_load_g2s(smem_raw + (size_t)sp * stage_bytes, cf + nxt * block_elems, block_bytes, &full[sp]);
            }
        }
        const double *start;
        Affine m0, m1c, m1r;      // shared coefficients: the maps of codes 0 (none), 1 (chemo), 2 (radio)
        if (!PER_T) {
            Affine m[4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
                m[a] = euler_affine(s_coef[4 * a], s_coef[4 * a + 1], s_coef[4 * a + 2], s_coef[4 * a + 3], u, h, substeps);
            m0 = m[0]; m1c = m[1]; m1r = m[2];
            // start value of (i,t) = xhat[t+1] = steps 0..t applied to F[0]: affine scan, lane owns steps 2 lane, 2 lane + 1
            // (and 64 + 2 lane, 65 + 2 lane for T > 65)
            Affine carry{1.0, 0.0};
            for (int base = 0; base < ns; base += 64) {
                const int k0 = base + 2 * lane;
                const Affine a0 = (k0 < ns) ? pick(m, option_to_code(s_code[b][k0])) : Affine{1.0, 0.0};
                const Affine a1 = (k0 + 1 < ns) ? pick(m, option_to_code(s_code[b][k0 + 1])) : Affine{1.0, 0.0};
                Affine inc = compose(a1, a0);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    Affine prev;
                    prev.al = __shfl_up_sync(0xffffffffu, inc.al, d);
                    prev.be = __shfl_up_sync(0xffffffffu, inc.be, d);
                    if (lane >= d) inc = compose(inc, prev);
                }
                Affine exc;
                exc.al = __shfl_up_sync(0xffffffffu, inc.al, 1);
                exc.be = __shfl_up_sync(0xffffffffu, inc.be, 1);
                if (lane == 0) exc = Affine{1.0, 0.0};
                const Affine upto0 = compose(compose(a0, exc), carry), upto1 = compose(inc, carry);
                if (k0 < ns) s_start[warp][k0] = apply(upto0, f0);
                if (k0 + 1 < ns) s_start[warp][k0 + 1] = apply(upto1, f0);
                Affine last;
                last.al = __shfl_sync(0xffffffffu, inc.al, 31);
                last.be = __shfl_sync(0xffffffffu, inc.be, 31);
                carry = compose(last, carry);
            }
            __syncwarp();
            start = s_start[warp];
        } else {
            if (tid < ns) {
                const int t = tid;
                const double *c = coefs + (i * W + t) * 16;
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    s_aff[t][a] = euler_affine(keep_term(c[4 * a], drop_below), keep_term(c[4 * a + 1], drop_below),
                                               keep_term(c[4 * a + 2], drop_below), keep_term(c[4 * a + 3], drop_below), u,
                                               h, substeps);
                const Affine *m = s_aff[t];
                double v = f0;
#pragma unroll 4
                for (int k = 0; k <= t; ++k) v = apply(m[option_to_code(s_code[b][k])], v);
                s_start[0][t] = v;
            }
            __syncthreads();
            start = s_start[0];
            m0 = m1c = m1r = Affine{1.0, 0.0};
        }
        mbar_wait(&full[s], parity);
        const double *buf = reinterpret_cast<const double *>(smem_raw + (size_t)s * stage_bytes);
        int t = tid / O, o = tid - t * O;
        const int dt_ = EV_THREADS / O, do_ = EV_THREADS - dt_ * O;
        for (int e = tid; e < npairs; e += EV_THREADS) {
            if (t < ns && ((s_valid[b][t] >> o) & 1)) {
                Affine b0 = m0, b1 = (o < H) ? m1c : m1r;   // sliding options: chemo (o < H) or radio once, at step o mod H
                if (PER_T) { b0 = s_aff[t][0]; b1 = s_aff[t][o < H ? 1 : 2]; }
                const int kk = o < H ? o : o - H;
                double v = start[t];
                const double *tg = buf + e * H;
#pragma unroll
                for (int k = 0; k < EV_MAXH; ++k) {
                    if (k < H) {
                        v = (k == kk) ? apply(b1, v) : apply(b0, v);
                        const double d = v - tg[k];
                        se[k] = fma(d, d, se[k]);
                    }
                }
                cnt += 1.0;
            }
            t += dt_; o += do_;
            if (o >= O) { o -= O; ++t; }
        }
    }
    // block combine, ordered grid combine
#pragma unroll
    for (int k = 0; k < EV_MAXH; ++k) {
        const double v = warp_sum(se[k]);
        if (lane == 0) s_red[warp][k] = v;
    }
    {
        const double v = warp_sum(cnt);
        if (lane == 0)
            for (int k = 0; k < EV_MAXH; ++k) s_red[warp][EV_MAXH + k] = v;
    }
    __syncthreads();
    constexpr int NV = 2 * EV_MAXH;
    if (tid < NV) {
        double v = 0.0;
        for (int w = 0; w < EV_WARPS; ++w) v += s_red[w][tid];
        partials[(size_t)blockIdx.x * NV + tid] = v;
    }
    if (grid_arrive(ticket, &s_flag)) {
        if (tid < NV) {
            double v = 0.0;
            for (unsigned int bb = 0; bb < gridDim.x; ++bb) v += __ldcg(&partials[(size_t)bb * NV + tid]);
            const int which = tid / EV_MAXH, k = tid % EV_MAXH;
            if (k < H) sums[which * H + k] = v;
        }
        if (tid == 0) *ticket = 0u;
    }
}

static int eval_scratch(void **scratch, size_t partial_bytes, cudaStream_t st)
{
    int rc = pool_alloc(scratch, 256 + partial_bytes, st);
    if (rc) return rc;
    rc = check_cuda(cudaMemsetAsync(*scratch, 0, 256, st), "cf_eval: memset ticket");
    if (rc) pool_free(*scratch, st);
    return rc;
}

}  // namespace b200i

using namespace b200i;

extern "C" int b200i_cf_eval_one_step(int64_t n, int32_t T, double dt, int32_t substeps, const double *factual,
                                      const uint8_t *codes, const double *cf, const int32_t *n_steps,
                                      const double *static_feature, const double *coefs, int32_t coefs_per_step,
                                      double drop_below, double *sums, void *stream)
{
    B200I_REQUIRE(n >= 0 && sums, B200I_E_ARG, "cf_eval_one_step: negative n or NULL sums");
    B200I_REQUIRE(T >= 3 && T <= EV_MAXT, B200I_E_UNSUPPORTED, "cf_eval_one_step: T=%d outside [3,%d]", T, EV_MAXT);
    B200I_REQUIRE(dt > 0 && substeps >= 1, B200I_E_ARG, "cf_eval_one_step: dt must be > 0 and substeps >= 1");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int W = T - 1;
    B200I_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (3 * W + 2), st));
    if (n == 0) return 0;
    B200I_REQUIRE(factual && codes && cf && n_steps && static_feature && coefs, B200I_E_ARG, "cf_eval_one_step: NULL argument");
    int64_t grid = (n + EV_WARPS - 1) / EV_WARPS;
    int64_t cap = (int64_t)num_sms() * 8;
    if (cap > EV_MAX_BLOCKS) cap = EV_MAX_BLOCKS;
    if (grid > cap) grid = cap;
    void *scratch = nullptr;
    int rc = eval_scratch(&scratch, (size_t)grid * 3 * EV_MAXT * sizeof(double), st);
    if (rc) return rc;
    unsigned int *ticket = static_cast<unsigned int *>(scratch);
    double *partials = reinterpret_cast<double *>(static_cast<uint8_t *>(scratch) + 256);
    const bool small = T <= 64;
    const void *kern = coefs_per_step ? (small ? reinterpret_cast<const void *>(cf_eval_one_step_kernel<true, 64>)
                                               : reinterpret_cast<const void *>(cf_eval_one_step_kernel<true, 128>))
                                      : (small ? reinterpret_cast<const void *>(cf_eval_one_step_kernel<false, 64>)
                                               : reinterpret_cast<const void *>(cf_eval_one_step_kernel<false, 128>));
    double hh = dt / substeps;
    void *args[] = {&n, &T, &hh, &substeps, &factual, &codes, &cf, &n_steps, &static_feature, &coefs, &drop_below, &partials,
                    &ticket, &sums};
    rc = check_cuda(cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(EV_THREADS), args, 0, st), "cf_eval_one_step launch");
    pool_free(scratch, st);
    return rc;
}

extern "C" int b200i_cf_eval_treatment_seq(int64_t n, int32_t T, int32_t H, double dt, int32_t substeps,
                                           const double *factual, const uint8_t *codes, const double *cf,
                                           const uint16_t *valid, const int32_t *n_steps, const double *static_feature,
                                           const double *coefs, int32_t coefs_per_step, double drop_below, double *sums,
                                           void *stream)
{
    B200I_REQUIRE(n >= 0 && sums, B200I_E_ARG, "cf_eval_treatment_seq: negative n or NULL sums");
    B200I_REQUIRE(T >= 3 && T <= EV_MAXT && H >= 1 && H <= EV_MAXH, B200I_E_UNSUPPORTED,
                  "cf_eval_treatment_seq: T=%d outside [3,%d] or H=%d outside [1,%d]", T, EV_MAXT, H, EV_MAXH);
    B200I_REQUIRE(dt > 0 && substeps >= 1, B200I_E_ARG, "cf_eval_treatment_seq: dt must be > 0 and substeps >= 1");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B200I_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * H, st));
    if (n == 0) return 0;
    B200I_REQUIRE(factual && codes && cf && valid && n_steps && static_feature && coefs, B200I_E_ARG,
                  "cf_eval_treatment_seq: NULL argument");
    B200I_REQUIRE(aligned16(cf), B200I_E_ALIGN, "cf_eval_treatment_seq: cf must be 16-byte aligned (bulk copies)");
    const size_t block_bytes = (size_t)(T - 1) * 2 * H * H * sizeof(double);
    const size_t stage_bytes = (block_bytes + 127) & ~(size_t)127;
    const int smem = (int)(EV_STAGES * stage_bytes);
    B200I_REQUIRE(smem <= 200 * 1024, B200I_E_UNSUPPORTED, "cf_eval_treatment_seq: T=%d, H=%d need %d bytes of shared memory", T,
                  H, smem);
    const bool h5 = H == 5;     // projection_horizon of the reference's configuration (config/dataset/cancer_sim.yaml:16)
    const void *kern = coefs_per_step ? (h5 ? reinterpret_cast<const void *>(cf_eval_seq_kernel<true, 5>)
                                            : reinterpret_cast<const void *>(cf_eval_seq_kernel<true, 0>))
                                      : (h5 ? reinterpret_cast<const void *>(cf_eval_seq_kernel<false, 5>)
                                            : reinterpret_cast<const void *>(cf_eval_seq_kernel<false, 0>));
    int per_sm = 1;
    {
        int rc0 = ensure_dyn_smem(kern, smem, EV_THREADS, &per_sm);
        if (rc0) return rc0;
    }
    if (per_sm < 1) per_sm = 1;
    int64_t grid = n;
    int64_t cap = (int64_t)num_sms() * per_sm;
    if (cap > EV_MAX_BLOCKS) cap = EV_MAX_BLOCKS;
    if (grid > cap) grid = cap;
    void *scratch = nullptr;
    int rc = eval_scratch(&scratch, (size_t)grid * 2 * EV_MAXH * sizeof(double), st);
    if (rc) return rc;
    unsigned int *ticket = static_cast<unsigned int *>(scratch);
    double *partials = reinterpret_cast<double *>(static_cast<uint8_t *>(scratch) + 256);
    double hh = dt / substeps;
    void *args[] = {&n, &T, &H, &hh, &substeps, &factual, &codes, &cf, &valid, &n_steps, &static_feature, &coefs, &drop_below,
                    &partials, &ticket, &sums};
    rc = check_cuda(cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(EV_THREADS), args, (size_t)smem, st),
                    "cf_eval_treatment_seq launch");
    if (rc) { pool_free(scratch, st); return rc; }
    rc = check_cuda(cudaGetLastError(), "cf_eval_treatment_seq launch");
    pool_free(scratch, st);
    return rc;
}
