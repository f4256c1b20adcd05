// sim_factual.cu -- K1: simulate_factual (cancer_simulation.py:218-375, loop :282-354) on sm_100a.
//
// One thread per patient walks the T columns sequentially.  Kernels, newest first:
//   * sim_factual_ws (sim_factual_ws.cuh, the default): in-place tiles of {16 columns x 32 patients} per array moved
//     by TMA, software-pipelined lean column arithmetic (fastmath.cuh), per-tile fallback to the generic column
//     function; optional row pitch; optional fused population statistics.
//   * sim_factual_tma<P,TC,SIN,GRAM> (first generation, variants 2-9): persistent CTAs of P threads, separate input
//     and output tiles, library log/exp/cbrt.  Kept as an independent cross-check of the lean kernel.
//   * sim_factual_generic: the same per-column arithmetic with direct global accesses; covers odd T, unaligned
//     buffers, the `assigned_actions` fixed policy and the last n % 128 rows of the row-class mapping.
// With GRAM the population statistics of theta_gram (K4) are accumulated while simulating.
#include "fastmath.cuh"
#include "philox.cuh"
#include "sim_math.cuh"
#include "stats_reduce.cuh"
#include "tma.cuh"
#include <type_traits>

namespace b200i {

struct SimC {
    double death, density, sphere, chemo_amt, radio_amt, decay, fd_dt, inv_sphere;
    int window;
};

struct FactualState {
    double V, C, D;     // column t-1
    double win[15];     // diameters of the last `window` volumes, oldest first
    int cnt;
    int code_prev;      // treatment code (chemo + 2*radio) of column t-1
    int t_end;          // last simulated column
    bool exists, alive;
};

struct Column {
    double V, C, D, ca, ra, pc, pr, death, recov;
};

struct Moments {
    double sv, svv, sc, scc, sd, sdd;
    __device__ __forceinline__ void clear() { sv = svv = sc = scc = sd = sdd = 0.0; }
    __device__ __forceinline__ void add(double v, double c, double d)
    {
        sv += v; svv += v * v; sc += c; scc += c * c; sd += d; sdd += d * d;
    }
};

__device__ __forceinline__ void state_init(FactualState &s, bool exists)
{
    s.V = s.C = s.D = 0.0;
#pragma unroll
    for (int j = 0; j < 15; ++j) s.win[j] = 0.0;
    s.cnt = 0; s.code_prev = 0; s.t_end = 0;
    s.exists = exists; s.alive = exists;
}

// produces column t of all nine outputs for one patient and advances the state
// FULL: caller guarantees t > 0 and a completely filled 15-slot window (steady state, t >= 16)
template <bool GRAM, bool FULL = false>
__device__ __forceinline__ void factual_column(int t, int T, const SimC &c, const Patient &p, FactualState &s,
                                               double noise, double urec, double uchemo, double uradio,
                                               const double *assigned, Column &o, PatientGram &pg, Moments &mom)
{
    o.V = o.C = o.D = o.ca = o.ra = o.pc = o.pr = o.death = o.recov = 0.0;
    if (!FULL && t == 0) {
        if (s.exists) {
            o.V = p.v0;
            s.V = p.v0;
            if (GRAM) mom.add(p.v0, 0.0, 0.0);
        }
        return;
    }
    if (!s.alive || t >= T - 1) return;

    double Vn = gompertz_step(p, s.V, s.C, s.D, noise);
    double metric;
    if (FULL) {
#pragma unroll
        for (int j = 0; j < 14; ++j) s.win[j] = s.win[j + 1];
        s.win[14] = calc_diameter(s.V, c.sphere);
        metric = np_mean_full(s.win);
    } else {
        window_push(s.win, s.cnt, c.window, calc_diameter(s.V, c.sphere));
        metric = np_mean(s.win, s.cnt);
    }
    double pr, pc;
    if (assigned != nullptr) {
        pc = assigned[0];
        pr = assigned[1];
    } else {
        pr = sigmoid_prob(p.radio_beta, metric, p.radio_int);
        pc = p.same_sigmoid ? pr : sigmoid_prob(p.chemo_beta, metric, p.chemo_int);
    }
    const bool ra = uradio < pr;
    const bool ca = uchemo < pc;
    const double D = ra ? c.radio_amt : 0.0;
    const double C = __dadd_rn(__dmul_rn(s.C, c.decay), ca ? c.chemo_amt : 0.0);
    const bool death = Vn > c.death;
    if (death) Vn = c.death;
    const bool recov = !death && recovery_test<true>(urec, Vn, c.density);
    if (recov) Vn = 0.0;
    const int code = (ca ? 1 : 0) + (ra ? 2 : 0);
    if (GRAM) {
        // sample k = t-1 of the SINDy regression (pkpd/utils.py:433-462 + FiniteDifference order 1)
        const double xdot = __ddiv_rn(__dsub_rn(Vn, s.V), c.fd_dt);
        pg.add(s.code_prev, s.V, xdot);
        if (code != s.code_prev) pg.add(s.code_prev, Vn, xdot);  // snippet end-point: backward difference
        mom.add(Vn, C, D);
    }
    o.V = Vn; o.C = C; o.D = D;
    o.ca = ca ? 1.0 : 0.0; o.ra = ra ? 1.0 : 0.0;
    o.pc = pc; o.pr = pr;
    o.death = death ? 1.0 : 0.0; o.recov = recov ? 1.0 : 0.0;
    s.V = Vn; s.C = C; s.D = D;
    s.code_prev = code;
    s.t_end = t;
    if (death || recov) s.alive = false;
}

// last regression sample of a patient: x[L] is the never-simulated zero at column seq_len
template <bool GRAM>
__device__ __forceinline__ void factual_finish(const SimC &c, const FactualState &s, PatientGram &pg)
{
    if (GRAM && s.exists) {
        const double xdot = __ddiv_rn(__dsub_rn(0.0, s.V), c.fd_dt);
        pg.add(s.code_prev, s.V, xdot);
        pg.add(s.code_prev, 0.0, xdot);
    }
}

// folds one patient's statistics into the warp accumulators (all 32 lanes participate)
__device__ __forceinline__ void fold_patient_stats(double *warp_acc, int lane, const PatientGram &pg,
                                                   const Moments &mom, double u, bool exists, int seq_len)
{
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double g[B200I_GRAM_PER_TREATMENT];
        expand_gram(pg.s[a], u, g);
#pragma unroll
        for (int j = 0; j < B200I_GRAM_PER_TREATMENT; ++j)
            warp_acc_add(warp_acc, a * B200I_GRAM_PER_TREATMENT + j, exists ? g[j] : 0.0, lane);
    }
    const int m0 = 4 * B200I_GRAM_PER_TREATMENT;
    warp_acc_add(warp_acc, m0 + 0, exists ? mom.sv : 0.0, lane);
    warp_acc_add(warp_acc, m0 + 1, exists ? mom.svv : 0.0, lane);
    warp_acc_add(warp_acc, m0 + 2, exists ? mom.sc : 0.0, lane);
    warp_acc_add(warp_acc, m0 + 3, exists ? mom.scc : 0.0, lane);
    warp_acc_add(warp_acc, m0 + 4, exists ? mom.sd : 0.0, lane);
    warp_acc_add(warp_acc, m0 + 5, exists ? mom.sdd : 0.0, lane);
    warp_acc_add(warp_acc, m0 + 6, exists ? (double)seq_len : 0.0, lane);
    warp_acc_add(warp_acc, m0 + 7, exists ? 1.0 : 0.0, lane);
}

// ------------------------------------------------------------------------------------------------
// generic kernel: direct global accesses
// ------------------------------------------------------------------------------------------------
struct FactualPtrs {
    const double *noise, *rec, *chemo_rvs, *radio_rvs, *assigned;
    double *out[9];  // V C D ca ra pc pr death recov
    double *seq_len;
};

template <bool GRAM>
__global__ void __launch_bounds__(128) sim_factual_generic(int64_t n, int64_t pstride, int T, SimC c,
                                                           const double *__restrict__ params, FactualPtrs io,
                                                           const double *__restrict__ static_feature,
                                                           StatsWorkspace *ws)
{
    __shared__ double block_acc[STATS_MAX_WARPS][STATS_PAD];
    __shared__ unsigned int s_is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (GRAM) {
        for (int j = threadIdx.x; j < STATS_MAX_WARPS * STATS_PAD; j += blockDim.x) (&block_acc[0][0])[j] = 0.0;
        __syncthreads();
    }
    const int64_t span = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = ((n + blockDim.x - 1) / blockDim.x) * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += span) {
        const bool exists = i < n;
        Patient p = load_patient(params, pstride, exists ? i : 0);
        FactualState s;
        state_init(s, exists);
        PatientGram pg; Moments mom;
        pg.clear(); mom.clear();
        if (exists) {
            const int64_t row = i * T;
            for (int t = 0; t < T; ++t) {
                Column o;
                const double *aa = io.assigned ? io.assigned + (row + t) * 2 : nullptr;
                factual_column<GRAM>(t, T, c, p, s, io.noise[row + t], io.rec[row + t], io.chemo_rvs[row + t],
                                     io.radio_rvs[row + t], aa, o, pg, mom);
                io.out[0][row + t] = o.V; io.out[1][row + t] = o.C; io.out[2][row + t] = o.D;
                io.out[3][row + t] = o.ca; io.out[4][row + t] = o.ra; io.out[5][row + t] = o.pc;
                io.out[6][row + t] = o.pr; io.out[7][row + t] = o.death; io.out[8][row + t] = o.recov;
            }
            factual_finish<GRAM>(c, s, pg);
            io.seq_len[i] = (double)(s.t_end + 1);
        }
        if (GRAM) {
            const double u = exists ? static_feature[i] : 0.0;
            fold_patient_stats(block_acc[warp], lane, pg, mom, u, exists, s.t_end + 1);
        }
    }
    if (GRAM) stats_block_finish(block_acc, blockDim.x >> 5, ws, &s_is_last);
}

// ------------------------------------------------------------------------------------------------
// TMA-tiled kernel
// ------------------------------------------------------------------------------------------------
struct TmapPack {
    CUtensorMap in[4];   // noise, recovery, chemo_rvs, radio_rvs
    CUtensorMap out[9];  // V C D ca ra pc pr death recov
};

template <int P, int TC, int SIN>
struct TileCfg {
    static constexpr int ROW_BYTES = TC * 8;
    static constexpr int TILE_BYTES = P * ROW_BYTES;
    static constexpr int IN_BYTES = SIN * 4 * TILE_BYTES;
    static constexpr int OUT_BYTES = 9 * TILE_BYTES;
    static constexpr int SMEM_BYTES = IN_BYTES + OUT_BYTES + 1024;  // + alignment slack
    static_assert(TILE_BYTES % 1024 == 0, "tiles must keep 1024-byte alignment");
    static_assert(ROW_BYTES == 32 || ROW_BYTES == 64 || ROW_BYTES == 128, "row = swizzle span");
};

template <int P, int TC, int SIN, int MINB, bool GRAM>
__global__ void __launch_bounds__(P, MINB)
sim_factual_tma(const __grid_constant__ TmapPack maps, int64_t n, int T, SimC c, const double *__restrict__ params,
                double *__restrict__ seq_len_out, const double *__restrict__ static_feature, StatsWorkspace *ws)
{
    using Cfg = TileCfg<P, TC, SIN>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[SIN];
    __shared__ double block_acc[GRAM ? STATS_MAX_WARPS : 1][STATS_PAD];
    __shared__ unsigned int s_is_last;

    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *in_tiles = smem;                    // [SIN][4][TILE]
    uint8_t *out_tiles = smem + Cfg::IN_BYTES;   // [9][TILE]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunks = (T + TC - 1) / TC;
    const int64_t ntiles = (n + P - 1) / P;
    const int64_t my_tiles = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total = my_tiles * nchunks;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SIN; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
#pragma unroll
        for (int a = 0; a < 4; ++a) tma_prefetch_desc(&maps.in[a]);
#pragma unroll
        for (int a = 0; a < 9; ++a) tma_prefetch_desc(&maps.out[a]);
    }
    if (GRAM) {
        for (int j = tid; j < STATS_MAX_WARPS * STATS_PAD; j += P) (&block_acc[0][0])[j] = 0.0;
    }
    __syncthreads();

    auto issue_load = [&](int64_t g) {
        const int stage = (int)(g % SIN);
        const int64_t tile = blockIdx.x + (g / nchunks) * gridDim.x;
        const int ch = (int)(g % nchunks);
        mbar_arrive_expect_tx(&bars[stage], 4u * Cfg::TILE_BYTES);
#pragma unroll
        for (int a = 0; a < 4; ++a)
            tma_load_2d(in_tiles + (stage * 4 + a) * Cfg::TILE_BYTES, &maps.in[a], ch * TC, (int)(tile * P), &bars[stage]);
    };

    if (tid == 0) {
        for (int64_t g = 0; g < SIN && g < total; ++g) issue_load(g);
    }

    Patient p;
    FactualState s;
    PatientGram pg;
    Moments mom;
    int64_t patient = 0;

    for (int64_t g = 0; g < total; ++g) {
        const int stage = (int)(g % SIN);
        const uint32_t parity = (uint32_t)((g / SIN) & 1);
        const int64_t tile = blockIdx.x + (g / nchunks) * gridDim.x;
        const int ch = (int)(g % nchunks);
        if (ch == 0) {
            patient = tile * P + tid;
            const bool exists = patient < n;
            p = load_patient(params, n, exists ? patient : 0);
            state_init(s, exists);
            pg.clear(); mom.clear();
        }
        if (tid == 0) tma_store_wait_read();  // previous chunk's stores no longer read the out tiles
        __syncthreads();
        mbar_wait(&bars[stage], parity);

        const uint8_t *in_base = in_tiles + stage * 4 * Cfg::TILE_BYTES;
        auto run_chunk = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll 1
            for (int q = 0; q < TC / 2; ++q) {
                const uint32_t off = swz_off<Cfg::ROW_BYTES>(tid, q);
                const double2 nz = *reinterpret_cast<const double2 *>(in_base + 0 * Cfg::TILE_BYTES + off);
                const double2 ur = *reinterpret_cast<const double2 *>(in_base + 1 * Cfg::TILE_BYTES + off);
                const double2 uc = *reinterpret_cast<const double2 *>(in_base + 2 * Cfg::TILE_BYTES + off);
                const double2 ud = *reinterpret_cast<const double2 *>(in_base + 3 * Cfg::TILE_BYTES + off);
                Column o0, o1;
                const int t0 = ch * TC + 2 * q;
                factual_column<GRAM, FULL>(t0, T, c, p, s, nz.x, ur.x, uc.x, ud.x, nullptr, o0, pg, mom);
                factual_column<GRAM, FULL>(t0 + 1, T, c, p, s, nz.y, ur.y, uc.y, ud.y, nullptr, o1, pg, mom);
                *reinterpret_cast<double2 *>(out_tiles + 0 * Cfg::TILE_BYTES + off) = make_double2(o0.V, o1.V);
                *reinterpret_cast<double2 *>(out_tiles + 1 * Cfg::TILE_BYTES + off) = make_double2(o0.C, o1.C);
                *reinterpret_cast<double2 *>(out_tiles + 2 * Cfg::TILE_BYTES + off) = make_double2(o0.D, o1.D);
                *reinterpret_cast<double2 *>(out_tiles + 3 * Cfg::TILE_BYTES + off) = make_double2(o0.ca, o1.ca);
                *reinterpret_cast<double2 *>(out_tiles + 4 * Cfg::TILE_BYTES + off) = make_double2(o0.ra, o1.ra);
                *reinterpret_cast<double2 *>(out_tiles + 5 * Cfg::TILE_BYTES + off) = make_double2(o0.pc, o1.pc);
                *reinterpret_cast<double2 *>(out_tiles + 6 * Cfg::TILE_BYTES + off) = make_double2(o0.pr, o1.pr);
                *reinterpret_cast<double2 *>(out_tiles + 7 * Cfg::TILE_BYTES + off) = make_double2(o0.death, o1.death);
                *reinterpret_cast<double2 *>(out_tiles + 8 * Cfg::TILE_BYTES + off) = make_double2(o0.recov, o1.recov);
            }
        };
        // the 15-slot diameter window is full from column 16 on: whole chunks take the steady-state path
        if (ch * TC >= 16 && c.window == 15)
            run_chunk(std::true_type{});
        else
            run_chunk(std::false_type{});
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int a = 0; a < 9; ++a)
                tma_store_2d(&maps.out[a], ch * TC, (int)(tile * P), out_tiles + a * Cfg::TILE_BYTES);
            tma_store_commit();
            if (g + SIN < total) issue_load(g + SIN);
        }
        if (ch == nchunks - 1) {
            factual_finish<GRAM>(c, s, pg);
            if (s.exists) seq_len_out[patient] = (double)(s.t_end + 1);
            if (GRAM) {
                const double u = s.exists ? __ldg(static_feature + patient) : 0.0;
                fold_patient_stats(block_acc[warp], lane, pg, mom, u, s.exists, s.t_end + 1);
            }
        }
    }
    if (tid == 0) tma_store_wait_all();
    if (GRAM) stats_block_finish(block_acc, P >> 5, ws, &s_is_last);
}

}  // namespace b200i
#include "sim_factual_ws.cuh"
#include "sim_factual_rng.cuh"
namespace b200i {

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <int P, int TC, int SIN, int MINB, bool GRAM>
static int launch_tma(int64_t n, int T, const SimC &c, const double *params, const double *const in[4],
                      double *const out[9], double *seq_len, const double *static_feature, StatsWorkspace *ws,
                      cudaStream_t st)
{
    using Cfg = TileCfg<P, TC, SIN>;
    TmapPack pack;
    for (int a = 0; a < 4; ++a) {
        int rc = encode_tmap_2d_f64(&pack.in[a], in[a], (uint64_t)n, (uint64_t)T, P, TC, true);
        if (rc) return rc;
    }
    for (int a = 0; a < 9; ++a) {
        int rc = encode_tmap_2d_f64(&pack.out[a], out[a], (uint64_t)n, (uint64_t)T, P, TC, false);
        if (rc) return rc;
    }
    auto kern = sim_factual_tma<P, TC, SIN, MINB, GRAM>;
    B200I_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int per_sm = 0;
    B200I_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, P, Cfg::SMEM_BYTES));
    B200I_REQUIRE(per_sm >= 1, B200I_E_UNSUPPORTED, "sim_factual_tma<%d,%d,%d>: does not fit on an SM", P, TC, SIN);
    const int64_t ntiles = (n + P - 1) / P;
    int64_t grid = (int64_t)num_sms() * per_sm;
    if (grid > ntiles) grid = ntiles;
    if (grid > STATS_MAX_BLOCKS) grid = STATS_MAX_BLOCKS;
    kern<<<(unsigned)grid, P, Cfg::SMEM_BYTES, st>>>(pack, n, T, c, params, seq_len, static_feature, ws);
    return check_cuda(cudaGetLastError(), "sim_factual_tma launch");
}

struct SideOut {
    uint8_t *codes;
    int64_t code_pitch;
    double *pmom;
};

template <bool GRAM>
static int dispatch_tma(int variant, int64_t n, int T, int64_t pitch, const SimC &c, const double *params,
                        const double *const in[4], double *const out[9], double *seq_len, const double *sf,
                        StatsWorkspace *ws, cudaStream_t st, SideOut side = SideOut{nullptr, 0, nullptr})
{
    if (side.codes != nullptr) {   // side outputs for the lean fit: the two default shapes only
        B200I_REQUIRE(!GRAM && (variant == 10 || variant == 12), B200I_E_UNSUPPORTED,
                      "sim_factual: side outputs need variant 0, 10 or 12 without the fused statistics");
        if (variant == 10)
            return launch_ws<32, 2, 6, 0, false, 2>(n, n, T, pitch, c, params, in, out, seq_len, st, nullptr, nullptr,
                                                    side.codes, side.code_pitch, side.pmom);
        return launch_ws<32, 1, 11, 0, false, 2>(n, n, T, pitch, c, params, in, out, seq_len, st, nullptr, nullptr,
                                                 side.codes, side.code_pitch, side.pmom);
    }
    B200I_REQUIRE(pitch == T || variant == 10 || variant == 12 || variant == 20 || variant == 21, B200I_E_UNSUPPORTED,
                  "sim_factual: variant %d needs dense rows (row_pitch %lld, T %d)", variant, (long long)pitch, T);
    switch (variant) {
        case 2: return launch_tma<128, 8, 1, 2, GRAM>(n, T, c, params, in, out, seq_len, sf, ws, st);
        // generation 6 (sim_factual_ws.cuh): <patients per tile, 16-column boxes per chunk, CTAs per SM, mode>;
        // 2x = data movement only (profiling aid).  The other tile shapes of generations 1-5 and the line-aligned
        // row-class mapping (variants 3-9, 11, 13-17: parity-green but slower, profiles/r1_k1_sweep_gen2.json,
        // r1_k1_skeleton.md) were removed in round 2; variant 2 stays as the first-generation cross-check.
        case 10: return launch_ws<32, 2, 6, 0, false, GRAM ? 1 : 0>(n, n, T, pitch, c, params, in, out, seq_len, st, sf, ws);
        case 12: return launch_ws<32, 1, 11, 0, false, GRAM ? 1 : 0>(n, n, T, pitch, c, params, in, out, seq_len, st, sf, ws);
        case 20: if (!GRAM) return launch_ws<32, 2, 6, 1, false>(n, n, T, pitch, c, params, in, out, seq_len, st); break;
        case 21: if (!GRAM) return launch_ws<32, 1, 11, 1, false>(n, n, T, pitch, c, params, in, out, seq_len, st); break;
        default:
            break;
    }
    set_error("sim_factual: unknown variant %d (fused gram=%d)", variant, (int)GRAM);
    return B200I_E_ARG;
}

}  // namespace b200i

using namespace b200i;

extern "C" int64_t b200i_gram_workspace_bytes(void) { return (int64_t)sizeof(StatsWorkspace); }

extern "C" int b200i_sim_factual(int64_t n, int32_t T, const b200i_sim_consts *k, const double *params,
                                 const double *noise, const double *recovery_rvs, const double *chemo_rvs,
                                 const double *radio_rvs, const double *assigned_actions, double *cancer_volume,
                                 double *chemo_dosage, double *radio_dosage, double *chemo_application,
                                 double *radio_application, double *chemo_probabilities, double *radio_probabilities,
                                 double *death_flags, double *recovery_flags, double *sequence_lengths,
                                 const double *static_feature, double fd_dt, void *gram_workspace, int32_t variant,
                                 void *stream)
{
    return b200i_sim_factual_pitched(n, T, T, k, params, noise, recovery_rvs, chemo_rvs, radio_rvs, assigned_actions,
                                     cancer_volume, chemo_dosage, radio_dosage, chemo_application, radio_application,
                                     chemo_probabilities, radio_probabilities, death_flags, recovery_flags,
                                     sequence_lengths, static_feature, fd_dt, gram_workspace, variant, stream);
}

static int sim_factual_impl(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *k, const double *params,
                            const double *noise, const double *recovery_rvs, const double *chemo_rvs,
                            const double *radio_rvs, const double *assigned_actions, double *cancer_volume,
                            double *chemo_dosage, double *radio_dosage, double *chemo_application,
                            double *radio_application, double *chemo_probabilities, double *radio_probabilities,
                            double *death_flags, double *recovery_flags, double *sequence_lengths,
                            const double *static_feature, double fd_dt, void *gram_workspace, int32_t variant,
                            void *stream, SideOut side);

extern "C" int b200i_sim_factual_side(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *k,
                                      const double *params, const double *noise, const double *recovery_rvs,
                                      const double *chemo_rvs, const double *radio_rvs, double *cancer_volume,
                                      double *chemo_dosage, double *radio_dosage, double *chemo_application,
                                      double *radio_application, double *chemo_probabilities,
                                      double *radio_probabilities, double *death_flags, double *recovery_flags,
                                      double *sequence_lengths, uint8_t *codes_out, int64_t code_pitch,
                                      double *patient_moments_out, int32_t variant, void *stream)
{
    B200I_REQUIRE(codes_out && patient_moments_out, B200I_E_ARG, "sim_factual_side: codes_out / patient_moments_out missing");
    // the kernel stores the code bytes of every 16-column box (the ragged last one included) as one 16-byte word
    B200I_REQUIRE(code_pitch % 16 == 0 && code_pitch >= ((T + 15) / 16) * 16, B200I_E_ARG,
                  "sim_factual_side: code_pitch %lld (multiple of 16, >= T rounded up to 16)", (long long)code_pitch);
    B200I_REQUIRE(aligned16(codes_out), B200I_E_ALIGN, "sim_factual_side: codes_out must be 16-byte aligned");
    return sim_factual_impl(n, T, row_pitch, k, params, noise, recovery_rvs, chemo_rvs, radio_rvs, nullptr, cancer_volume,
                            chemo_dosage, radio_dosage, chemo_application, radio_application, chemo_probabilities,
                            radio_probabilities, death_flags, recovery_flags, sequence_lengths, nullptr, 1.0, nullptr,
                            variant, stream, SideOut{codes_out, code_pitch, patient_moments_out});
}

extern "C" int b200i_sim_factual_pitched(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *k,
                                         const double *params, const double *noise, const double *recovery_rvs,
                                         const double *chemo_rvs, const double *radio_rvs,
                                         const double *assigned_actions, double *cancer_volume, double *chemo_dosage,
                                         double *radio_dosage, double *chemo_application, double *radio_application,
                                         double *chemo_probabilities, double *radio_probabilities, double *death_flags,
                                         double *recovery_flags, double *sequence_lengths,
                                         const double *static_feature, double fd_dt, void *gram_workspace,
                                         int32_t variant, void *stream)
{
    return sim_factual_impl(n, T, row_pitch, k, params, noise, recovery_rvs, chemo_rvs, radio_rvs, assigned_actions,
                            cancer_volume, chemo_dosage, radio_dosage, chemo_application, radio_application,
                            chemo_probabilities, radio_probabilities, death_flags, recovery_flags, sequence_lengths,
                            static_feature, fd_dt, gram_workspace, variant, stream, SideOut{nullptr, 0, nullptr});
}

static int sim_factual_impl(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *k, const double *params,
                            const double *noise, const double *recovery_rvs, const double *chemo_rvs,
                            const double *radio_rvs, const double *assigned_actions, double *cancer_volume,
                            double *chemo_dosage, double *radio_dosage, double *chemo_application,
                            double *radio_application, double *chemo_probabilities, double *radio_probabilities,
                            double *death_flags, double *recovery_flags, double *sequence_lengths,
                            const double *static_feature, double fd_dt, void *gram_workspace, int32_t variant,
                            void *stream, SideOut side)
{
    const int64_t pitch = row_pitch;
    B200I_REQUIRE(pitch == T || (pitch > T && pitch % 2 == 0), B200I_E_ARG, "sim_factual: row_pitch %lld (T = %d) must be even and >= T",
                  (long long)pitch, T);
    B200I_REQUIRE(n >= 0, B200I_E_ARG, "sim_factual: negative n");
    if (n == 0) return 0;
    B200I_REQUIRE(k && params && noise && recovery_rvs && chemo_rvs && radio_rvs && cancer_volume &&
                      chemo_dosage && radio_dosage && chemo_application && radio_application && chemo_probabilities &&
                      radio_probabilities && death_flags && recovery_flags && sequence_lengths,
                  B200I_E_ARG, "sim_factual: NULL argument");
    B200I_REQUIRE(T >= 3 && T <= 4096, B200I_E_UNSUPPORTED, "sim_factual: seq_length %d outside [3,4096]", T);
    B200I_REQUIRE(k->lag == 0, B200I_E_UNSUPPORTED, "sim_factual: lag=%d (only lag=0 is implemented)", k->lag);
    B200I_REQUIRE(k->window_size >= 1 && k->window_size <= 15, B200I_E_UNSUPPORTED,
                  "sim_factual: window_size=%d outside [1,15]", k->window_size);
    B200I_REQUIRE(n < (int64_t)1 << 31, B200I_E_UNSUPPORTED, "sim_factual: n=%lld >= 2^31", (long long)n);
    const bool gram = gram_workspace != nullptr;
    B200I_REQUIRE(!gram || (static_feature && fd_dt > 0), B200I_E_ARG,
                  "sim_factual: fused gram needs static_feature and fd_dt > 0");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SimC c{k->death_threshold, k->cell_density, k->sphere_coef, k->chemo_amt, k->radio_amt, k->drug_decay, fd_dt,
           1.0 / k->sphere_coef, k->window_size};
    const double *in[4] = {noise, recovery_rvs, chemo_rvs, radio_rvs};
    double *out[9] = {cancer_volume, chemo_dosage, radio_dosage, chemo_application, radio_application,
                      chemo_probabilities, radio_probabilities, death_flags, recovery_flags};
    StatsWorkspace *ws = static_cast<StatsWorkspace *>(gram_workspace);

    bool tma_ok = (T % 2 == 0) && assigned_actions == nullptr;
    for (int a = 0; a < 4; ++a) tma_ok = tma_ok && aligned16(in[a]);
    for (int a = 0; a < 9; ++a) tma_ok = tma_ok && aligned16(out[a]);
    // auto: rows on 128-byte lines -> one box per chunk and 11 warps per SM; otherwise two boxes per chunk
    if (variant == 0) variant = !tma_ok ? 1 : ((pitch * 8) % 128 == 0 ? 12 : 10);
    B200I_REQUIRE(pitch == T || variant >= 2, B200I_E_UNSUPPORTED, "sim_factual: pitched rows need even T and 16-byte aligned arrays");
    if (variant >= 2) {
        B200I_REQUIRE(assigned_actions == nullptr, B200I_E_UNSUPPORTED,
                      "sim_factual: assigned_actions is only handled by the generic kernel (variant 1)");
        B200I_REQUIRE(tma_ok, B200I_E_ALIGN, "sim_factual: TMA variant needs even T and 16-byte aligned arrays");
        return gram ? dispatch_tma<true>(variant, n, T, pitch, c, params, in, out, sequence_lengths, static_feature, ws, st)
                    : dispatch_tma<false>(variant, n, T, pitch, c, params, in, out, sequence_lengths, static_feature, ws, st,
                                          side);
    }
    B200I_REQUIRE(side.codes == nullptr, B200I_E_UNSUPPORTED,
                  "sim_factual: side outputs need the tiled kernel (even T, 16-byte aligned arrays)");
    FactualPtrs io;
    io.noise = noise; io.rec = recovery_rvs; io.chemo_rvs = chemo_rvs; io.radio_rvs = radio_rvs;
    io.assigned = assigned_actions;
    for (int a = 0; a < 9; ++a) io.out[a] = out[a];
    io.seq_len = sequence_lengths;
    int64_t grid = (n + 127) / 128;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (grid > cap) grid = cap;
    if (gram)
        sim_factual_generic<true><<<(unsigned)grid, 128, 0, st>>>(n, n, T, c, params, io, static_feature, ws);
    else
        sim_factual_generic<false><<<(unsigned)grid, 128, 0, st>>>(n, n, T, c, params, io, static_feature, ws);
    return check_cuda(cudaGetLastError(), "sim_factual_generic launch");
}

// ------------------------------------------------------------------------------------------------
// K1L: device-generated draws (sim_factual_rng.cuh)
// ------------------------------------------------------------------------------------------------
extern "C" int b200i_philox_draws(int64_t n, int32_t T, int64_t row_pitch, uint64_t seed, int64_t patient_base,
                                  double *noise, double *recovery_rvs, double *chemo_rvs, double *radio_rvs,
                                  void *stream)
{
    B200I_REQUIRE(n >= 0 && patient_base >= 0, B200I_E_ARG, "philox_draws: negative n or patient_base");
    if (n == 0) return 0;
    B200I_REQUIRE(noise && recovery_rvs && chemo_rvs && radio_rvs, B200I_E_ARG, "philox_draws: NULL argument");
    B200I_REQUIRE(T >= 2 && T % 2 == 0 && row_pitch >= T && row_pitch % 2 == 0, B200I_E_UNSUPPORTED,
                  "philox_draws: T=%d and row_pitch=%lld must be even, row_pitch >= T", T, (long long)row_pitch);
    B200I_REQUIRE(aligned16(noise) && aligned16(recovery_rvs) && aligned16(chemo_rvs) && aligned16(radio_rvs),
                  B200I_E_ALIGN, "philox_draws: arrays must be 16-byte aligned");
    const int64_t total = n * (T / 2);
    int64_t grid = (total + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (grid > cap) grid = cap;
    philox_draws_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n, T, row_pitch, rng::round_keys((uint32_t)seed, (uint32_t)(seed >> 32)), (uint32_t)seed, (uint32_t)(seed >> 32),
        patient_base, noise, recovery_rvs, chemo_rvs, radio_rvs);
    return check_cuda(cudaGetLastError(), "philox_draws launch");
}

extern "C" int b200i_sim_factual_rng(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *k,
                                     const double *params, int64_t params_stride, uint64_t seed, int64_t patient_base,
                                     double *cancer_volume, uint8_t *codes_out, int64_t code_pitch,
                                     double *sequence_lengths, double *patient_moments_out, int64_t moments_stride,
                                     const double *static_feature, double fd_dt, void *gram_workspace, int32_t variant,
                                     void *stream)
{
    B200I_REQUIRE(n >= 0 && patient_base >= 0, B200I_E_ARG, "sim_factual_rng: negative n or patient_base");
    if (n == 0) {
        if (gram_workspace)
            B200I_CUDA(cudaMemsetAsync(gram_workspace, 0, sizeof(double) * 128 + sizeof(unsigned int) * 32,
                                       static_cast<cudaStream_t>(stream)));
        return 0;
    }
    B200I_REQUIRE(k && params && cancer_volume && sequence_lengths, B200I_E_ARG, "sim_factual_rng: NULL argument");
    B200I_REQUIRE(params_stride >= n && (patient_moments_out == nullptr || moments_stride >= n), B200I_E_ARG,
                  "sim_factual_rng: params_stride %lld / moments_stride %lld must be >= n", (long long)params_stride,
                  (long long)moments_stride);
    B200I_REQUIRE(T >= 4 && T <= 1024 && T % 2 == 0, B200I_E_UNSUPPORTED, "sim_factual_rng: seq_length %d (even, 4..1024)", T);
    B200I_REQUIRE(row_pitch >= T && row_pitch % 2 == 0 && aligned16(cancer_volume), B200I_E_ALIGN,
                  "sim_factual_rng: row_pitch %lld must be even and >= T, cancer_volume 16-byte aligned", (long long)row_pitch);
    B200I_REQUIRE(codes_out == nullptr || (code_pitch >= ((T + 15) / 16) * 16 && code_pitch % 16 == 0 && aligned16(codes_out)),
                  B200I_E_ARG, "sim_factual_rng: code_pitch %lld (multiple of 16, >= T rounded up to 16)", (long long)code_pitch);
    B200I_REQUIRE(k->lag == 0 && k->window_size >= 1 && k->window_size <= 15, B200I_E_UNSUPPORTED,
                  "sim_factual_rng: lag=%d window_size=%d (lag 0, window 1..15)", k->lag, k->window_size);
    B200I_REQUIRE(n < (int64_t)1 << 31, B200I_E_UNSUPPORTED, "sim_factual_rng: n=%lld >= 2^31", (long long)n);
    const bool gram = gram_workspace != nullptr;
    B200I_REQUIRE(!gram || (static_feature && fd_dt > 0), B200I_E_ARG,
                  "sim_factual_rng: fused statistics need static_feature and fd_dt > 0");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SimC c{k->death_threshold, k->cell_density, k->sphere_coef, k->chemo_amt, k->radio_amt, k->drug_decay,
           fd_dt > 0 ? fd_dt : 1.0, 1.0 / k->sphere_coef, k->window_size};
    CUtensorMap vmap;
    {
        int rc = encode_tmap_2d_pitched_f64(&vmap, cancer_volume, (uint64_t)n, (uint64_t)T, (uint64_t)row_pitch * 8, 32, 16);
        if (rc) return rc;
    }
    // variant 0 / 2: second generation (phased, one column per loop body, 16 warps per SM); 1: first generation
    // (four unrolled columns with the generator inlined, 12 warps per SM) -- kept as an independent cross-check
    B200I_REQUIRE(variant >= 0 && variant <= 2, B200I_E_UNSUPPORTED, "sim_factual_rng: variant %d (0 auto, 1, 2)", variant);
    StatsWorkspace *ws = static_cast<StatsWorkspace *>(gram_workspace);
    if (gram) {
        B200I_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 128 + sizeof(unsigned int) * 32, st));
        if (variant == 1)
            return launch_rng<1, 3, 1>(vmap, n, params_stride, 0, T, c, params, seed, patient_base, codes_out, code_pitch,
                                       sequence_lengths, nullptr, static_feature, ws, st);
        return launch_rng<1, 3, 2>(vmap, n, params_stride, 0, T, c, params, seed, patient_base, codes_out, code_pitch,
                                   sequence_lengths, nullptr, static_feature, ws, st);
    }
    if (patient_moments_out) {
        if (variant == 1)
            return launch_rng<2, 3, 1>(vmap, n, params_stride, moments_stride, T, c, params, seed, patient_base, codes_out,
                                       code_pitch, sequence_lengths, patient_moments_out, nullptr, nullptr, st);
        return launch_rng<2, 4, 2>(vmap, n, params_stride, moments_stride, T, c, params, seed, patient_base, codes_out,
                                   code_pitch, sequence_lengths, patient_moments_out, nullptr, nullptr, st);
    }
    if (variant == 1)
        return launch_rng<0, 3, 1>(vmap, n, params_stride, 0, T, c, params, seed, patient_base, codes_out, code_pitch,
                                   sequence_lengths, nullptr, nullptr, nullptr, st);
    return launch_rng<0, 4, 2>(vmap, n, params_stride, 0, T, c, params, seed, patient_base, codes_out, code_pitch,
                               sequence_lengths, nullptr, nullptr, nullptr, st);
}

// Host parameters -> device, chunk by chunk on `copy_stream`, each chunk simulated on `stream` as soon as it has
// arrived (events), so the PCIe transfer of chunk c+1 overlaps the simulation of chunk c.  One call replaces
// ~12 framework calls per chunk of the Python pipeline (the step is ~2 ms: host-side launch cost matters).
// sum of the per-chunk statistics in chunk order (fixed order => reproducible bits for a given chunk count)
__global__ void sum_chunk_stats_kernel(const uint8_t *ws_base, size_t ws_stride, int chunks, double *out)
{
    const int i = threadIdx.x;
    if (i >= B200I_STATS_DOUBLES) return;
    double v = 0.0;
    for (int c = 0; c < chunks; ++c) v += reinterpret_cast<const StatsWorkspace *>(ws_base + (size_t)c * ws_stride)->stats[i];
    out[i] = v;
}

// parameter rows that are one scalar for the whole cohort are filled on the device instead of being copied
struct UniformRows {
    double v[B200I_NUM_PARAMS];
};
__global__ void __launch_bounds__(256)
fill_uniform_rows_kernel(double *__restrict__ params, int64_t n, uint32_t mask, UniformRows u)
{
    for (int r = 0; r < B200I_NUM_PARAMS; ++r) {
        if (!((mask >> r) & 1u)) continue;
        const double v = u.v[r];
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
            params[r * n + i] = v;
    }
}

// rows of a chunk that are derived on the device instead of crossing PCIe: beta = alpha / 10 (get_standard_params,
// cancer_simulation.py:185: the same IEEE division) and the static feature from the patient type byte
__global__ void __launch_bounds__(256)
derive_chunk_rows_kernel(double *__restrict__ params, int64_t n, int64_t a, int64_t count, int derive_beta,
                         const uint8_t *__restrict__ types_u8, double *__restrict__ static_feature)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        if (derive_beta) params[3 * n + a + i] = __ddiv_rn(params[1 * n + a + i], 10.0);
        if (types_u8) static_feature[a + i] = (double)types_u8[a + i];
    }
}

static int upload_simulate_rng_impl(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *k,
                                    const double *params_host, uint32_t uniform_mask,
                                    const double *uniform_values_host, const double *static_host, double *params,
                                    double *static_feature, uint64_t seed, int64_t patient_base,
                                    double *cancer_volume, uint8_t *codes_out, int64_t code_pitch,
                                    double *sequence_lengths, double *patient_moments_out, int32_t chunks,
                                    double fd_dt, void *chunk_gram_workspaces, double *stats_out,
                                    void *copy_stream, void *stream, int derive_beta, const uint8_t *types_u8_host,
                                    uint8_t *types_u8_dev, int overlap_prev = 0);

extern "C" int b200i_upload_simulate_rng(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *k,
                                         const double *params_host, uint32_t uniform_mask,
                                         const double *uniform_values_host, const double *static_host, double *params,
                                         double *static_feature, uint64_t seed, int64_t patient_base,
                                         double *cancer_volume, uint8_t *codes_out, int64_t code_pitch,
                                         double *sequence_lengths, double *patient_moments_out, int32_t chunks,
                                         double fd_dt, void *chunk_gram_workspaces, double *stats_out,
                                         void *copy_stream, void *stream)
{
    return upload_simulate_rng_impl(n, T, row_pitch, k, params_host, uniform_mask, uniform_values_host, static_host, params,
                                    static_feature, seed, patient_base, cancer_volume, codes_out, code_pitch,
                                    sequence_lengths, patient_moments_out, chunks, fd_dt, chunk_gram_workspaces, stats_out,
                                    copy_stream, stream, 0, nullptr, nullptr);
}

extern "C" int b200i_upload_simulate_rng_reduced(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *k,
                                                 const double *params_host, uint32_t uniform_mask,
                                                 const double *uniform_values_host, int32_t derive_beta,
                                                 const uint8_t *patient_types_host, uint8_t *patient_types_dev,
                                                 double *params, double *static_feature, uint64_t seed,
                                                 int64_t patient_base, double *cancer_volume, uint8_t *codes_out,
                                                 int64_t code_pitch, double *sequence_lengths, double *patient_moments_out,
                                                 int32_t chunks, double fd_dt, void *chunk_gram_workspaces,
                                                 double *stats_out, void *copy_stream, void *stream)
{
    B200I_REQUIRE(patient_types_host && patient_types_dev && static_feature, B200I_E_ARG,
                  "upload_simulate_rng_reduced: patient_types_host / patient_types_dev / static_feature is NULL");
    B200I_REQUIRE(!derive_beta || !((uniform_mask >> 3) & 1u), B200I_E_ARG,
                  "upload_simulate_rng_reduced: beta cannot be both derived and uniform");
    return upload_simulate_rng_impl(n, T, row_pitch, k, params_host, uniform_mask, uniform_values_host, nullptr, params,
                                    static_feature, seed, patient_base, cancer_volume, codes_out, code_pitch,
                                    sequence_lengths, patient_moments_out, chunks, fd_dt, chunk_gram_workspaces, stats_out,
                                    copy_stream, stream, derive_beta ? 1 : 0, patient_types_host, patient_types_dev);
}

extern "C" int b200i_upload_simulate_rng_pipelined(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *k,
                                                   const double *params_host, uint32_t uniform_mask,
                                                   const double *uniform_values_host, int32_t derive_beta,
                                                   const uint8_t *patient_types_host, uint8_t *patient_types_dev,
                                                   double *params, double *static_feature, uint64_t seed,
                                                   int64_t patient_base, double *cancer_volume, uint8_t *codes_out,
                                                   int64_t code_pitch, double *sequence_lengths, double *patient_moments_out,
                                                   int32_t chunks, double fd_dt, void *chunk_gram_workspaces,
                                                   double *stats_out, void *copy_stream, void *stream)
{
    B200I_REQUIRE(patient_types_host && patient_types_dev && static_feature, B200I_E_ARG,
                  "upload_simulate_rng_pipelined: patient_types_host / patient_types_dev / static_feature is NULL");
    B200I_REQUIRE(!derive_beta || !((uniform_mask >> 3) & 1u), B200I_E_ARG,
                  "upload_simulate_rng_pipelined: beta cannot be both derived and uniform");
    return upload_simulate_rng_impl(n, T, row_pitch, k, params_host, uniform_mask, uniform_values_host, nullptr, params,
                                    static_feature, seed, patient_base, cancer_volume, codes_out, code_pitch,
                                    sequence_lengths, patient_moments_out, chunks, fd_dt, chunk_gram_workspaces, stats_out,
                                    copy_stream, stream, derive_beta ? 1 : 0, patient_types_host, patient_types_dev, 1);
}

static int upload_simulate_rng_impl(int64_t n, int32_t T, int64_t row_pitch, const b200i_sim_consts *k,
                                    const double *params_host, uint32_t uniform_mask,
                                    const double *uniform_values_host, const double *static_host, double *params,
                                    double *static_feature, uint64_t seed, int64_t patient_base,
                                    double *cancer_volume, uint8_t *codes_out, int64_t code_pitch,
                                    double *sequence_lengths, double *patient_moments_out, int32_t chunks,
                                    double fd_dt, void *chunk_gram_workspaces, double *stats_out,
                                    void *copy_stream, void *stream, int derive_beta, const uint8_t *types_u8_host,
                                    uint8_t *types_u8_dev, int overlap_prev)
{
    const bool fit = chunk_gram_workspaces != nullptr;
    B200I_REQUIRE(!fit || (stats_out && static_feature && codes_out && patient_moments_out && fd_dt > 0), B200I_E_ARG,
                  "upload_simulate_rng: the per-chunk fit needs stats_out, static_feature, codes_out, patient_moments_out, fd_dt");
    const size_t ws_stride = (size_t)b200i_gram_workspace_bytes();
    B200I_REQUIRE(n >= 0 && chunks >= 1 && chunks <= 64, B200I_E_ARG, "upload_simulate_rng: n=%lld chunks=%d (1..64)",
                  (long long)n, chunks);
    if (n == 0) return 0;
    B200I_REQUIRE(params_host && params && copy_stream && copy_stream != stream, B200I_E_ARG,
                  "upload_simulate_rng: NULL argument, or copy_stream == stream (no overlap possible)");
    B200I_REQUIRE(types_u8_host != nullptr || (static_host == nullptr) == (static_feature == nullptr), B200I_E_ARG,
                  "upload_simulate_rng: static_host and static_feature must both be given or both be NULL");
    B200I_REQUIRE(uniform_mask < (1u << B200I_NUM_PARAMS) && (uniform_mask == 0 || uniform_values_host), B200I_E_ARG,
                  "upload_simulate_rng: uniform_mask 0x%x needs uniform_values_host[10]", uniform_mask);
    cudaStream_t cs = static_cast<cudaStream_t>(copy_stream), st = static_cast<cudaStream_t>(stream);
    int64_t step = (n + chunks - 1) / chunks;
    step = ((step + 31) / 32) * 32;   // whole 32-patient tiles per chunk
    // chunk kernels alternate between `stream` and an internal second stream, so that the next chunk fills the SMs
    // the previous one is draining (each chunk is only one or two waves of 32-patient tiles)
    static thread_local cudaStream_t aux[16] = {};
    int devid = 0;
    B200I_CUDA(cudaGetDevice(&devid));
    B200I_REQUIRE(devid >= 0 && devid < 16, B200I_E_UNSUPPORTED, "upload_simulate_rng: device index %d", devid);
    if (aux[devid] == nullptr) B200I_CUDA(cudaStreamCreateWithFlags(&aux[devid], cudaStreamNonBlocking));
    cudaStream_t sx = aux[devid];
    if (uniform_mask) {   // stream-ordered after the previous readers of the block, before every chunk (event below)
        UniformRows u;
        for (int r = 0; r < B200I_NUM_PARAMS; ++r) u.v[r] = uniform_values_host[r];
        fill_uniform_rows_kernel<<<num_sms() * 4, 256, 0, st>>>(params, n, uniform_mask, u);
        B200I_CUDA(cudaGetLastError());
    }
    // the previous work on `stream` may still read the parameter block / write the outputs.  One event per (thread,
    // device), created once: the call creates and destroys nothing, so it can be captured into a CUDA graph
    // (cohort.GeneratedFitPipeline.step_host(graph=True): one graph launch instead of ~60 stream operations per step)
    static thread_local cudaEvent_t evs[16] = {};
    if (evs[devid] == nullptr) B200I_CUDA(cudaEventCreateWithFlags(&evs[devid], cudaEventDisableTiming));
    cudaEvent_t ev = evs[devid];
    // overlap_prev (b200i_upload_simulate_rng_pipelined): the copies of this call do not wait for everything queued on
    // `stream` before it (the previous step's all-reduce, STLSQ and result copies) but, chunk by chunk, only for the
    // kernels of the previous call that read that chunk's parameter rows -- the upload of step s+1 then runs under the
    // tail of step s.  The kernels themselves stay ordered behind `stream` as before.
    static thread_local cudaEvent_t chunk_done[16][64] = {};
    int rc = check_cuda(cudaEventRecord(ev, st), "cudaEventRecord");
    if (!rc && !overlap_prev) rc = check_cuda(cudaStreamWaitEvent(cs, ev, 0), "cudaStreamWaitEvent");
    if (!rc) rc = check_cuda(cudaStreamWaitEvent(sx, ev, 0), "cudaStreamWaitEvent");
    int c = 0;
    for (int64_t a = 0; a < n && !rc; a += step, ++c) {
        const int64_t b = (a + step < n) ? a + step : n;
        cudaStream_t run = (c & 1) ? sx : st;
        if (overlap_prev && chunk_done[devid][c]) rc = check_cuda(cudaStreamWaitEvent(cs, chunk_done[devid][c], 0), "cudaStreamWaitEvent");
        if (rc) break;
        const uint32_t skip = uniform_mask | (derive_beta ? (1u << 3) : 0u);   // rows that do not cross PCIe
        for (int r0 = 0; r0 < B200I_NUM_PARAMS && !rc;) {   // maximal runs of rows that are real arrays
            if ((skip >> r0) & 1u) { ++r0; continue; }
            int r1 = r0;
            while (r1 < B200I_NUM_PARAMS && !((skip >> r1) & 1u)) ++r1;
            rc = check_cuda(cudaMemcpy2DAsync(params + (size_t)r0 * n + a, (size_t)n * 8, params_host + (size_t)r0 * n + a,
                                              (size_t)n * 8, (size_t)(b - a) * 8, r1 - r0, cudaMemcpyHostToDevice, cs),
                            "cudaMemcpy2DAsync(params)");
            r0 = r1;
        }
        if (!rc && static_host)
            rc = check_cuda(cudaMemcpyAsync(static_feature + a, static_host + a, (size_t)(b - a) * 8,
                                            cudaMemcpyHostToDevice, cs), "cudaMemcpyAsync(static)");
        if (!rc && types_u8_host)
            rc = check_cuda(cudaMemcpyAsync(types_u8_dev + a, types_u8_host + a, (size_t)(b - a), cudaMemcpyHostToDevice, cs),
                            "cudaMemcpyAsync(patient types)");
        if (!rc) rc = check_cuda(cudaEventRecord(ev, cs), "cudaEventRecord");
        if (!rc) rc = check_cuda(cudaStreamWaitEvent(run, ev, 0), "cudaStreamWaitEvent");
        if (!rc && (derive_beta || types_u8_host)) {
            int64_t g = (b - a + 255) / 256;
            if (g > num_sms() * 4) g = num_sms() * 4;
            derive_chunk_rows_kernel<<<(unsigned)g, 256, 0, run>>>(params, n, a, b - a, derive_beta, types_u8_dev, static_feature);
            rc = check_cuda(cudaGetLastError(), "derive_chunk_rows launch");
        }
        if (!rc)
            rc = b200i_sim_factual_rng(b - a, T, row_pitch, k, params + a, n, seed, patient_base + a,
                                       cancer_volume + a * row_pitch, codes_out ? codes_out + a * code_pitch : nullptr,
                                       code_pitch, sequence_lengths + a, patient_moments_out ? patient_moments_out + a : nullptr,
                                       n, nullptr, 0.0, nullptr, 0, run);
        if (!rc && fit)   // the chunk's share of the population statistics, right behind its simulation
            rc = b200i_theta_gram_codes(b - a, T, row_pitch, 0, fd_dt, cancer_volume + a * row_pitch, codes_out + a * code_pitch,
                                        code_pitch, sequence_lengths + a, static_feature + a, patient_moments_out + a, n,
                                        static_cast<uint8_t *>(chunk_gram_workspaces) + (size_t)c * ws_stride, run);
        if (!rc && overlap_prev) {   // the readers of this chunk's rows are queued: the next call's copy of the chunk waits for them
            if (chunk_done[devid][c] == nullptr)
                rc = check_cuda(cudaEventCreateWithFlags(&chunk_done[devid][c], cudaEventDisableTiming), "cudaEventCreate");
            if (!rc) rc = check_cuda(cudaEventRecord(chunk_done[devid][c], run), "cudaEventRecord");
        }
    }
    if (!rc && c > 1) {   // `stream` continues after the chunks of the second stream
        rc = check_cuda(cudaEventRecord(ev, sx), "cudaEventRecord");
        if (!rc) rc = check_cuda(cudaStreamWaitEvent(st, ev, 0), "cudaStreamWaitEvent");
    }
    if (!rc && fit) {
        sum_chunk_stats_kernel<<<1, 96, 0, st>>>(static_cast<const uint8_t *>(chunk_gram_workspaces), ws_stride, c, stats_out);
        rc = check_cuda(cudaGetLastError(), "sum_chunk_stats launch");
    }
    return rc;
}
