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
// bulk copy (cp.async.bulk, SASS UBLKCP) into a 3-stage shared-memory ring.
#include "stats_reduce.cuh"
#include "tma.cuh"

namespace b200i {

constexpr int EV_THREADS = 128;
constexpr int EV_WARPS = EV_THREADS / 32;
constexpr int EV_STAGES = 3;
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
template <bool PER_T>
__global__ void __launch_bounds__(EV_THREADS)
cf_eval_one_step_kernel(int64_t n, int T, double h, int substeps, const double *__restrict__ F,
                        const uint8_t *__restrict__ codes, const double *__restrict__ cf,
                        const int *__restrict__ n_steps, const double *__restrict__ static_u,
                        const double *__restrict__ coefs, double drop_below, double *__restrict__ partials,
                        unsigned int *__restrict__ ticket, double *__restrict__ sums)
{
    __shared__ double s_F[EV_WARPS][EV_MAXT];
    __shared__ __align__(16) double s_cf[EV_WARPS][EV_MAXT * 4];
    __shared__ uint8_t s_code[EV_WARPS][EV_MAXT];
    __shared__ double s_coef[16];
    __shared__ double s_acc[EV_WARPS][3 * EV_MAXT];
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
    for (int64_t i = (int64_t)blockIdx.x * EV_WARPS + warp; i < n; i += nwarps) {
        int ns = n_steps[i];
        if (ns > W) ns = W;
        const double u = static_u[i];
        __syncwarp();
        for (int k = lane; k < T; k += 32) {
            s_F[warp][k] = F[i * T + k];
            s_code[warp][k] = codes[i * T + k];
        }
        for (int e = lane; e < 4 * W; e += 32) s_cf[warp][e] = cf[i * (int64_t)(4 * W) + e];
        __syncwarp();
        const double F0 = s_F[warp][0];
        if (!PER_T) {
            Affine m[4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
                m[a] = euler_affine(s_coef[4 * a], s_coef[4 * a + 1], s_coef[4 * a + 2], s_coef[4 * a + 3], u, h, substeps);
            // affine prefix scan over the factual steps: lane owns steps 4 lane .. 4 lane + 3
            Affine comp{1.0, 0.0};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = 4 * lane + q;
                if (k < ns) comp = compose(pick(m, option_to_code(s_code[warp][k])), comp);
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
                        const double p = apply(pick(m, option_to_code(o)), v);
                        const double target = (o == fo) ? s_F[warp][k + 1] : s_cf[warp][4 * k + o];
                        const double d = p - target;
                        if (o == fo) e_f = d * d;
                        else e_o += d * d;
                    }
                    // rows (i, t > k) all carry the factual step k: 4 (ns-1-k) rows; the 4 rows of t = k end here
                    se[q] += e_f * (double)(4 * (ns - 1 - k) + 1) + e_o;
                    cnt[q] += (double)(4 * (ns - k));
                    last[q] += e_f + e_o;
                    v = apply(pick(m, option_to_code(fo)), v);
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
                            const double target = (o == fo) ? s_F[warp][k + 1] : s_cf[warp][4 * k + o];
                            const double d = pr - target;
                            e_all += d * d;
                        }
#pragma unroll
                        for (int j = 0; j < EV_SL; ++j)
                            if (j == p) { se[j] += e_all; last[j] += e_all; }
                    }
                    v = apply(pick(m, option_to_code(fo)), v);
                    const double d = v - s_F[warp][k + 1];
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
// ------------------------------------------------------------------------------------------------
template <bool PER_T>
__global__ void __launch_bounds__(EV_THREADS)
cf_eval_seq_kernel(int64_t n, int T, int H, double h, int substeps, const double *__restrict__ F,
                   const uint8_t *__restrict__ codes, const double *__restrict__ cf, const uint16_t *__restrict__ valid,
                   const int *__restrict__ n_steps, const double *__restrict__ static_u, const double *__restrict__ coefs,
                   double drop_below, double *__restrict__ partials, unsigned int *__restrict__ ticket,
                   double *__restrict__ sums)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t full[EV_STAGES];
    __shared__ Affine s_aff[PER_T ? EV_MAXT : 1][4];
    __shared__ double s_start[EV_MAXT];
    __shared__ uint8_t s_code[EV_MAXT];
    __shared__ uint16_t s_valid[EV_MAXT];
    __shared__ double s_coef[16];
    __shared__ double s_red[EV_WARPS][2 * EV_MAXH];
    __shared__ unsigned int s_flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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

    uint32_t it = 0;
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x, ++it) {
        const int s = (int)(it % EV_STAGES);
        const uint32_t parity = (it / EV_STAGES) & 1u;
        int ns = n_steps[i];
        if (ns > W) ns = W;
        const double u = static_u[i];
        if (tid < T) s_code[tid] = codes[i * T + tid];
        if (tid < W) s_valid[tid] = valid[i * W + tid];
        if (!PER_T && tid < 4)
            s_aff[0][tid] = euler_affine(s_coef[4 * tid], s_coef[4 * tid + 1], s_coef[4 * tid + 2], s_coef[4 * tid + 3], u, h,
                                         substeps);
        __syncthreads();
        // start value of (i,t): xhat[t+1], rolled from F[0] over the factual codes 0..t with the maps of (i,t)
        if (tid < ns) {
            const int t = tid;
            if (PER_T) {
                const double *c = coefs + (i * W + t) * 16;
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    s_aff[t][a] = euler_affine(keep_term(c[4 * a], drop_below), keep_term(c[4 * a + 1], drop_below),
                                               keep_term(c[4 * a + 2], drop_below), keep_term(c[4 * a + 3], drop_below), u,
                                               h, substeps);
            }
            const Affine *m = PER_T ? s_aff[t] : s_aff[0];
            double v = F[i * T];
            for (int k = 0; k <= t; ++k) v = apply(m[option_to_code(s_code[k])], v);
            s_start[t] = v;
        }
        __syncthreads();
        mbar_wait(&full[s], parity);
        const double *buf = reinterpret_cast<const double *>(smem_raw + (size_t)s * stage_bytes);
        for (int e = tid; e < npairs; e += EV_THREADS) {
            const int t = e / O, o = e - t * O;
            if (t < ns && ((s_valid[t] >> o) & 1)) {
                const Affine *m = PER_T ? s_aff[t] : s_aff[0];
                const Affine m0 = m[0], m1 = m[o < H ? 1 : 2];   // sliding options: chemo (o < H) or radio once, at step o mod H
                const int kk = o < H ? o : o - H;
                double v = s_start[t];
#pragma unroll
                for (int k = 0; k < EV_MAXH; ++k) {
                    if (k < H) {
                        v = apply(k == kk ? m1 : m0, v);
                        const double d = v - buf[e * H + k];
                        se[k] += d * d;
                    }
                }
                cnt += 1.0;
            }
        }
        __syncthreads();   // every thread is done with stage s, s_start and s_aff
        if (tid == 0) {
            const int64_t nxt = i + (int64_t)EV_STAGES * gridDim.x;
            if (nxt < n) {
                mbar_arrive_expect_tx(&full[s], block_bytes);
                bulk_load_g2s(smem_raw + (size_t)s * stage_bytes, cf + nxt * block_elems, block_bytes, &full[s]);
            }
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
            for (unsigned int b = 0; b < gridDim.x; ++b) v += __ldcg(&partials[(size_t)b * NV + tid]);
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
    if (coefs_per_step)
        cf_eval_one_step_kernel<true><<<(unsigned)grid, EV_THREADS, 0, st>>>(n, T, dt / substeps, substeps, factual, codes, cf,
                                                                              n_steps, static_feature, coefs, drop_below,
                                                                              partials, ticket, sums);
    else
        cf_eval_one_step_kernel<false><<<(unsigned)grid, EV_THREADS, 0, st>>>(n, T, dt / substeps, substeps, factual, codes, cf,
                                                                               n_steps, static_feature, coefs, drop_below,
                                                                               partials, ticket, sums);
    rc = check_cuda(cudaGetLastError(), "cf_eval_one_step launch");
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
    const void *kern = coefs_per_step ? reinterpret_cast<const void *>(cf_eval_seq_kernel<true>)
                                      : reinterpret_cast<const void *>(cf_eval_seq_kernel<false>);
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
    if (coefs_per_step)
        cf_eval_seq_kernel<true><<<(unsigned)grid, EV_THREADS, smem, st>>>(n, T, H, dt / substeps, substeps, factual, codes, cf,
                                                                           valid, n_steps, static_feature, coefs, drop_below,
                                                                           partials, ticket, sums);
    else
        cf_eval_seq_kernel<false><<<(unsigned)grid, EV_THREADS, smem, st>>>(n, T, H, dt / substeps, substeps, factual, codes,
                                                                            cf, valid, n_steps, static_feature, coefs,
                                                                            drop_below, partials, ticket, sums);
    rc = check_cuda(cudaGetLastError(), "cf_eval_treatment_seq launch");
    pool_free(scratch, st);
    return rc;
}
