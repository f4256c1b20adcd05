// sim_factual_ws.cuh -- K1 generation 6 (default): software-pipelined columns on in-place 128-byte-row tiles.
// Included by sim_factual.cu (needs TmapPack, SimC, FactualState/factual_column, tma.cuh, fastmath.cuh).
//
// What the measurements said about the earlier generations (B200, 1M patients x 60 columns):
//   * gen 1 ncu (profiles/r1_k1_gen1_variant2_ncu.txt): issue-bound on a bloated instruction stream --
//     489 warp instructions per patient-column, 25 % of them FP64, one dependent chain per column.
//   * a data-movement-only build of gen 4/5 (64-byte rows, then 128-byte rows requested one 16-column
//     chunk at a time) ran almost as slowly as the full kernel: short row segments that are touched ~5 us
//     apart defeat DRAM page locality and L2 line reuse (dram read 1.7x the algorithmic bytes, 50 % of
//     peak).  The same skeleton moving whole 480-byte rows reached 5.86 TB/s (profiles/r1_k1_skeleton.md).
//   * FP64 dependent-issue latency is 8 cycles, MUFU.RCP64H 17, I2F+D2I 36 (scripts/dbg/fp64_lat.cu): a warp
//     that walks one dependent chain per column cannot fill its issue slots, and shared memory (in-place
//     tiles cost 32 B per patient-column) caps the number of resident warps.
// Generation 6:
//   * a chunk is NB adjacent boxes of {16 columns x P patients} per array, all requested back to back, so
//     DRAM sees 128*NB contiguous bytes per row at once; every box is its own SWIZZLE_128B tile
//     (conflict-free 16-byte LDS/STS for thread-per-patient access).  One in-place stage per CTA: volume,
//     chemo dosage and the treatment probability overwrite the random draws and leave through TMA; the five
//     0/1-valued outputs travel as one packed byte per (patient, column) and are expanded to float64 with
//     coalesced 16-byte stores (radio dosage = dose x application).  Small single-stage CTAs (P = 32: one
//     warp, no CTA-wide barrier) overlap each other's copy phases.
//   * the column recurrence is software-pipelined into three independent chains per loop body:
//       (1) log(K / V[t-1])                                   -> V[t]
//       (2) exp / reciprocal of the sigmoid of column t-1     -> treatment, chemo concentration of t-1
//       (3) cube root of V[t-1], window mean                  -> sigmoid argument of column t
//     (2) consumes what (3) produced one body earlier; V[t] needs (1) and (2).  The critical path per column
//     drops from ~45 to ~26 dependent FP64 operations.
//   * arithmetic from fastmath.cuh with all polynomial constants pinned in registers; no special-case
//     paths in the hot loop.  A tile that does not satisfy the fast path's preconditions (chemo and radio
//     sigmoids differ, |sigmoid argument| could exceed 700, window_size != 15, non-normal K / V0) is
//     processed by the generic column function with the library log/exp/cbrt instead -- results stay
//     defined for any input.
//   * the 15-slot diameter window is an 18-register file shifted once per four columns (four columns are
//     unrolled so every window index is static); while it fills (t <= 15, exactly box 0 of chunk 0, which
//     has its own code copy) numpy's pairwise sum degenerates to a running sum plus one 8-leaf tree at t = 8.
//   * inactive columns (after death / recovery, and the never-simulated last column) are zeroed by a
//     rarely taken clean-up branch per four columns instead of predicating every output.
#pragma once

namespace b200i {

template <int P, int NB>
struct WsCfg {
    static constexpr int BOX_COLS = 16, TCH = BOX_COLS * NB;          // columns per chunk
    static constexpr int TILE_BYTES = P * 128;                        // one box of one array
    static constexpr int STAGE_BYTES = NB * 4 * TILE_BYTES;           // [NB][4][TILE]
    static constexpr int FLAG_PITCH = TCH + 4;                        // byte c+4 = column c (4-byte aligned); byte 3 = dummy
    static constexpr int FLAG_BYTES = P * FLAG_PITCH;
    static constexpr int DUMMY_BYTES = P * 16;                        // per-thread scratch slot
    static constexpr int LUT_BYTES = 5 * 4 * 16;                      // flag expansion table [array][2 bits] -> double2
    static constexpr int WARP_REGION = (STAGE_BYTES + FLAG_BYTES + DUMMY_BYTES + 1023) & ~1023;   // one tile's buffers
    static constexpr int SMEM_BYTES = WARP_REGION + LUT_BYTES + 1024;
    static constexpr int SMEM_BYTES_4 = 4 * WARP_REGION + LUT_BYTES + 1024;   // row-class mapping: four warps
    static_assert(P % (TCH / 2) == 0, "flag expansion mapping");
    static_assert(P % 32 == 0 && TILE_BYTES % 1024 == 0, "swizzled tiles need 1024-byte alignment");
};

// simulator constants + fastmath constants, loaded once per thread and pinned in registers
struct WsK {
    fm::FmK f;
    double sphere, inv_sphere, decay, dose, death, ndensity, inv15;
};

struct WsPatient {
    fm::LogNum K;                       // carrying capacity, pre-split for log_ratio
    double rho, beta_c, rd, nb, si;     // rd = alpha*d + beta*d^2 at d = radio dose; nb = -beta_sigmoid
};

struct WsState {
    double V;          // V[t-1]
    double Cq;         // C[t-2]
    double zq;         // sigmoid argument of column t-1
    double ucp, udp;   // chemo / radio draws of column t-1
    double S;          // running window sum while the window fills
    unsigned flp;      // death / recovery bits of column t-1 (its chemo / radio bits are still pending)
    bool alive;
    int t_end;
};

__device__ __forceinline__ void pin(double &x) { asm volatile("" : "+d"(x)); }

// numpy pairwise sum of 8 values
__device__ __forceinline__ double ws_tree8(double a0, double a1, double a2, double a3, double a4, double a5,
                                           double a6, double a7)
{
    return __dadd_rn(__dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3)), __dadd_rn(__dadd_rn(a4, a5), __dadd_rn(a6, a7)));
}

__device__ __noinline__ bool ws_recovery_rare(double u, double x)
{
    // u < exp(x) outside the hot path's shortcut (x >= 0 is handled by the caller)       :346
    return u < exp(x);
}

// treatment of column t-1 from its sigmoid argument (chain 2): probability, assignment, chemo concentration
__device__ __forceinline__ void ws_treat(const WsK &k, const WsState &s, double &pr, bool &ra, bool &ca, double &C1)
{
    pr = fm::rcp_fast(__dadd_rn(1.0, fm::exp_fast(k.f, s.zq)));                          // :322-323
    ra = s.udp < pr;                                                                      // :328-335
    ca = s.ucp < pr;
    C1 = __dadd_rn(__dmul_rn(s.Cq, k.decay), ca ? k.dose : 0.0);                          // :338
}

// Recovery draw of a column.  The tiled kernel passes the value it loaded; the generator kernel passes a lazy draw:
// its recovery uniforms lie in (0,1) with a smallest value of 2^-53 > exp(-40), so the draw can only matter when
// -V * density > -40 (V < 6.9e-8, i.e. an eradicated tumour) and is generated in that rare branch only.
__device__ __forceinline__ double ws_ur_value(double u) { return u; }
__device__ __forceinline__ double ws_ur_value(const rng::LazyRecovery &u) { return u.get(); }
template <class U> struct WsUrLazy { static constexpr bool value = false; };
template <> struct WsUrLazy<rng::LazyRecovery> { static constexpr bool value = true; };

// One loop body: volume of column t, treatment of column t-1, sigmoid argument of column t.
// FILL: the window is still filling (t <= 15); J = position in the unrolled group of four.
template <bool FILL, int J, class UR>
__device__ __forceinline__ void ws_body(int t, int Tm1, const WsK &k, const WsPatient &p, WsState &s, double (&w)[18],
                                        double v0, double nz, const UR &ur_in, double uc, double ud, double &oV,
                                        double &oC, double &oP, unsigned &oF)
{
    if (FILL && J == 0 && t == 0) {
        // column 0: the initial volume, nothing else                                     :282-289
        oV = v0; oC = 0.0; oP = 0.0; oF = 0u;
        s.V = v0;
        return;
    }
    // ---- chain 2: treatment of column t-1 (column 0 has none) ----
    double pr, C1;
    bool ra, ca;
    if (FILL && J == 1 && t == 1) {
        pr = 0.0; C1 = 0.0; ra = ca = false;
    } else {
        ws_treat(k, s, pr, ra, ca, C1);
    }
    // ---- chain 3: diameter window -> sigmoid argument of column t.  w holds cube roots; the factor 2 of
    // calc_diameter commutes exactly with the sum and the division                       :309-313
    const double cn = fm::cbrt_fast(fm::div_small(s.V, k.sphere, k.inv_sphere));
    w[14 + J] = cn;
    double mean;
    if (FILL) {
        if (J == 0 && t == 8)
            s.S = ws_tree8(w[7], w[8], w[9], w[10], w[11], w[12], w[13], w[14]);
        else
            s.S = __dadd_rn(s.S, cn);
        mean = fm::div_small(s.S, (double)t, fm::kInvN[t & 15]);
    } else {
        double r = ws_tree8(w[J], w[J + 1], w[J + 2], w[J + 3], w[J + 4], w[J + 5], w[J + 6], w[J + 7]);
#pragma unroll
        for (int j = 8; j < 15; ++j) r = __dadd_rn(r, w[J + j]);
        mean = fm::div_small(r, 15.0, k.inv15);
    }
    const double z = __dmul_rn(p.nb, __dsub_rn(__dmul_rn(mean, 2.0), p.si));
    // ---- chain 1: V * (1 + rho*log(K/V) - beta_c*C - (alpha*d + beta*d^2) + noise)      :300-302
    double g1 = __dadd_rn(1.0, __dmul_rn(p.rho, fm::log_ratio(k.f, p.K, s.V)));
    g1 = __dsub_rn(g1, __dmul_rn(p.beta_c, C1));
    g1 = __dsub_rn(g1, ra ? p.rd : 0.0);
    g1 = __dadd_rn(g1, nz);
    double Vn = __dmul_rn(s.V, g1);
    const bool act = s.alive && (t < Tm1);
    const bool death = Vn > k.death;                                                       // :340-343
    Vn = death ? k.death : Vn;
    // recovery: u < exp(-V * density); exp is below 4.3e-18 unless V < 6.9e-8             :346-349
    const double x = __dmul_rn(Vn, k.ndensity);
    bool recov = false;
    if (WsUrLazy<UR>::value) {
        if (act && !(x <= -40.0)) {   // also taken for NaN, like the reference's comparison
            const double ur = ws_ur_value(ur_in);
            recov = (x >= 0.0) ? (x == x) : ((x > -40.0) ? (ur < fm::exp_fast(k.f, x)) : ws_recovery_rare(ur, x));
        }
    } else {
        const double ur = ws_ur_value(ur_in);
        if (act && !(x <= -40.0 && ur >= 1e-17))   // also taken for NaN, like the reference's comparison
            recov = (x >= 0.0) ? (x == x) : ((x > -40.0) ? (ur < fm::exp_fast(k.f, x)) : ws_recovery_rare(ur, x));
    }
    recov = recov && !death;
    Vn = recov ? 0.0 : Vn;
    // outputs: volume and death/recovery bits of column t; treatment of column t-1
    oV = Vn; oC = C1; oP = pr;
    oF = s.flp | (ca ? 1u : 0u) | (ra ? 2u : 0u);
    s.flp = (death ? 4u : 0u) | (recov ? 8u : 0u);
    s.V = Vn; s.Cq = C1; s.zq = z; s.ucp = uc; s.udp = ud;
    s.t_end = act ? t : s.t_end;
    s.alive = act && !(death || recov);
}

// ---- generic (slow) path for tiles outside the fast path's preconditions ----------------------
struct WsSlow {
    Patient p;
    FactualState s;
    PatientGram pg;
    Moments mom;
};

// one regression sample (x0 -> x1 under treatment a) of the SINDy fit plus, when the constant-treatment snippet
// ends at x1, its backward-difference end point (pkpd/utils.py:433-462 + FiniteDifference order 1): both rows
// share xdot and the treatment, so they are merged and filed with one-hot weights (branch-free, 20 FMA chains)
__device__ __forceinline__ void ws_gram_sample(PatientGram &pg, bool valid, bool end, double x0, double x1, unsigned a,
                                               double fd_dt, double inv_dt)
{
    const double xdot = fm::div_small(__dsub_rn(x1, x0), fd_dt, inv_dt);
    const double e = end ? 1.0 : 0.0;
    const double cnt = 1.0 + e, sx = fma(e, x1, x0), sxx = fma(e * x1, x1, x0 * x0);
    const double sd = cnt * xdot, sxd = sx * xdot;
#pragma unroll
    for (unsigned t = 0; t < 4; ++t) {
        const double wt = (valid && a == t) ? 1.0 : 0.0;
        pg.s[t][0] = fma(wt, cnt, pg.s[t][0]); pg.s[t][1] = fma(wt, sx, pg.s[t][1]);
        pg.s[t][2] = fma(wt, sxx, pg.s[t][2]); pg.s[t][3] = fma(wt, sd, pg.s[t][3]);
        pg.s[t][4] = fma(wt, sxd, pg.s[t][4]);
    }
}

template <int P, bool GRAM>
__device__ __noinline__ void ws_slow_chunk(uint8_t *stage, uint8_t *flags_s, int tid, int t_first, int ncols, int T,
                                           const SimC &c, WsSlow *st)
{
    PatientGram &pg = st->pg;
    Moments &mom = st->mom;
    for (int cidx = 0; cidx < ncols; ++cidx) {
        if (t_first + cidx < 0) continue;   // skewed first box: columns left of column 0
        const int box = cidx >> 4, q = (cidx & 15) >> 1, lohi = cidx & 1;
        uint8_t *base = stage + box * 4 * (P * 128) + swz_off<128>((uint32_t)tid, (uint32_t)q) + lohi * 8;
        double *pn = reinterpret_cast<double *>(base), *pu = reinterpret_cast<double *>(base + 1 * P * 128),
               *pc = reinterpret_cast<double *>(base + 2 * P * 128), *pr = reinterpret_cast<double *>(base + 3 * P * 128);
        Column o;
        factual_column<GRAM, false>(t_first + cidx, T, c, st->p, st->s, *pn, *pu, *pc, *pr, nullptr, o, pg, mom);
        *pn = o.V; *pu = o.C; *pc = o.pc; *pr = o.pr;
        flags_s[cidx + 4] = (uint8_t)((o.ca != 0.0 ? 1u : 0u) | (o.ra != 0.0 ? 2u : 0u) | (o.death != 0.0 ? 4u : 0u) |
                                      (o.recov != 0.0 ? 8u : 0u));
    }
}

// opts bit 1: evict-first hint on every output store; bit 2: evict-last hint on the draw loads.
//
// MODE 0: simulator.  MODE 1: data movement only (threads copy draws to outputs) -- profiling aid that
// measures what the load/store skeleton sustains without the arithmetic.
//
// SKEW (the default): line-aligned row segments for T*8 not a multiple of 128.
//   The skeleton measurements (profiles/r1_k1_skeleton.md) show that what costs DRAM efficiency is not the
//   length of a row segment but its alignment: with 512-byte rows (T = 64) even 128-byte boxes sustain 5.7 TB/s,
//   with 480-byte rows (T = 60) the same boxes start on 32-byte boundaries, every request straddles two
//   128-byte lines that are fetched / written twice, and the skeleton drops to 3.5-4.5 TB/s.  For T = 60 four
//   consecutive rows are exactly 15 lines, and row j of such a group becomes line-aligned when its boxes start
//   at column 16 m - s_j with s_j = (j T) mod 16 (8 j T + 8 (16 m - s_j) = 0 mod 128).  So a work item is
//   (128 consecutive rows, class j): the 32 rows = j mod 4.  TMA rejects negative start coordinates, so each
//   class has its own 2-D maps (row pitch 4 T * 8 bytes) whose base is moved instead: the input view of class j
//   starts s_j columns before its first row (those columns alias the tail of the previous row: read, never
//   used), the output view starts at the first line boundary inside the row, and the row's first 16 - s_j
//   columns -- the partial line it shares with the previous row -- are written with ordinary 16-byte stores.
//   DRAM also wants the accesses of one moment to be dense in address space (giving each class to its own
//   CTA measured 1.6-1.9 ms: aligned, but every request then opens its own DRAM page).  So a CTA is four warps
//   = the four classes of one 128-row tile; each warp has its own tiles, barrier and column range (no
//   divergence inside a warp), and the four leaders issue their loads right after a CTA-wide barrier, so that
//   the memory system sees one line of each of 128 consecutive rows at once.  Rows beyond the last full
//   128-row tile go through the generic kernel.
// tensor maps of the row-class mapping: [class][array]; out arrays = volume, chemo dosage, chemo / radio probability
struct TmapPackSkew {
    CUtensorMap in[4][4];
    CUtensorMap out[4][4];
};
template <bool SKEW> struct WsMaps { typedef TmapPack type; };
template <> struct WsMaps<true> { typedef TmapPackSkew type; };

__device__ __forceinline__ const CUtensorMap *ws_in_map(const TmapPack &m, int, int a) { return &m.in[a]; }
__device__ __forceinline__ const CUtensorMap *ws_in_map(const TmapPackSkew &m, int j, int a) { return &m.in[j][a]; }
// o: 0 volume, 1 chemo dosage, 2 chemo probability, 3 radio probability
__device__ __forceinline__ const CUtensorMap *ws_out_map(const TmapPack &m, int, int o)
{
    return &m.out[o == 0 ? 0 : (o == 1 ? 1 : (o == 2 ? 5 : 6))];
}
__device__ __forceinline__ const CUtensorMap *ws_out_map(const TmapPackSkew &m, int j, int o) { return &m.out[j][o]; }

// STATS: 0 none; 1 fused population statistics (GRAM); 2 side outputs for the lean fit (SIDE): one treatment-code
// byte per (patient, column) and six per-patient moment sums, from which theta_gram_codes finishes the statistics
// reading 0.6 instead of 2.4 GB per million patients.
template <int P, int NB, int MINB, int MODE, bool SKEW, int STATS>
__global__ void __launch_bounds__(SKEW ? 128 : P, MINB)
sim_factual_ws(const __grid_constant__ typename WsMaps<SKEW>::type maps, int opts, int64_t n, int64_t pstride, int T,
               int64_t pitch, SimC c,
               const double *__restrict__ params, double *__restrict__ out_ca, double *__restrict__ out_ra,
               double *__restrict__ out_D, double *__restrict__ out_death, double *__restrict__ out_recov,
               double *__restrict__ seq_len_out, double *__restrict__ out_V, double *__restrict__ out_C,
               double *__restrict__ out_pc, double *__restrict__ out_pr, const double *__restrict__ static_feature,
               StatsWorkspace *ws, uint8_t *__restrict__ codes_out, int64_t code_pitch, double *__restrict__ pmom_out)
{
    constexpr bool GRAM = STATS == 1, SIDE = STATS == 2;
    static_assert(STATS == 0 || (!SKEW && MODE == 0), "statistics: plain tiles, simulator mode");
    __shared__ double block_acc[GRAM ? (P / 32) : 1][STATS_PAD];
    __shared__ unsigned int s_is_last;
    using Cfg = WsCfg<P, NB>;
    constexpr int TCH = Cfg::TCH, HALF = TCH / 2;
    static_assert(!SKEW || P == 32, "the row-class mapping is one warp per class");
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bars[4];
    const int tid = threadIdx.x;
    const int warp = SKEW ? (tid >> 5) : 0;            // row class in the row-class mapping
    const int rid = SKEW ? (tid & 31) : tid;           // row of the (warp's) tile
    const bool leader = rid == 0;
    uint64_t &full_bar = full_bars[warp];
    uint8_t *smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *tiles = smem_al + warp * Cfg::WARP_REGION;                              // [NB][4][TILE]
    uint8_t *flag_buf = tiles + Cfg::STAGE_BYTES;                                    // [P][FLAG_PITCH]
    uint8_t *dummy_buf = flag_buf + Cfg::FLAG_BYTES;                                 // [P][16]
    double2 *lut = reinterpret_cast<double2 *>(smem_al + (SKEW ? 4 : 1) * Cfg::WARP_REGION);   // [5][4]

    // work items: plain = tile of P rows; row classes = tile of 128 rows (warp j takes the rows = j mod 4)
    constexpr int ncls = 4;
    const int64_t n_items = SKEW ? n / 128 : (n + P - 1) / P;
    const bool opt_evict_first = (opts & 2) != 0, opt_evict_last = (opts & 4) != 0;

    struct Item {
        int64_t tile;   // first row = tile * (SKEW ? 128 : P)
        int j, shift, nboxes, nch;
    };
    auto item_of = [&](int64_t w) {
        Item it;
        if (SKEW) {
            it.tile = w; it.j = warp;
            it.shift = (it.j * T) & 15;                    // 8 (16 m - shift) + 8 j T = 0 mod 128 (T % 4 == 0)
        } else {
            it.tile = w; it.j = 0; it.shift = 0;
        }
        it.nboxes = (T + it.shift + 15) / 16;
        it.nch = (it.nboxes + NB - 1) / NB;
        return it;
    };
    auto cta_sync = [&]() {   // everybody who shares this tile
        if (P == 32) __syncwarp(); else __syncthreads();
    };
    // chunks per item: the row-class warps run in lockstep, so all of them loop over the longest class
    const int nch_all = SKEW ? ((T + 12 + 15) / 16 + NB - 1) / NB : ((T + 15) / 16 + NB - 1) / NB;
    // box m of the item: tile columns [16 m, 16 m + 16) = global columns [16 m - shift, ...)
    auto load_box = [&](void *dst, int a, const Item &it, int m, uint64_t pol) {
        const CUtensorMap *map = ws_in_map(maps, it.j, a);
        const int r0 = SKEW ? (int)(it.tile * 32) : (int)(it.tile * P);
        if (opt_evict_last) tma_load_2d_hint(dst, map, 16 * m, r0, &full_bar, pol);
        else tma_load_2d(dst, map, 16 * m, r0, &full_bar);
    };
    // the output views of a shifted class start at its second box (the first, partial one is stored by hand)
    auto store_box = [&](const void *src, int o, const Item &it, int m, uint64_t pol) {
        const CUtensorMap *map = ws_out_map(maps, it.j, o);
        const int r0 = SKEW ? (int)(it.tile * 32) : (int)(it.tile * P);
        const int c0 = 16 * (m - ((SKEW && it.shift > 0) ? 1 : 0));
        if (opt_evict_first) tma_store_2d_hint(map, c0, r0, src, pol);
        else tma_store_2d(map, c0, r0, src);
    };
    auto issue_load = [&](const Item &it, int ch) {
        int nb = it.nboxes - ch * NB;
        nb = nb > NB ? NB : nb;
        mbar_arrive_expect_tx(&full_bar, (uint32_t)(nb * 4 * Cfg::TILE_BYTES));
        const uint64_t pol = opt_evict_last ? l2_policy_evict_last() : 0ull;
        for (int b = 0; b < nb; ++b)
#pragma unroll
            for (int a = 0; a < 4; ++a)
                load_box(tiles + (b * 4 + a) * Cfg::TILE_BYTES, a, it, ch * NB + b, pol);
    };

    if (leader) {
        mbar_init(&full_bar, 1);
        mbar_fence_init();
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            tma_prefetch_desc(ws_in_map(maps, warp, a));
            tma_prefetch_desc(ws_out_map(maps, warp, a));
        }
    }
    __syncthreads();
    if (leader && (int64_t)blockIdx.x < n_items) issue_load(item_of(blockIdx.x), 0);
    if (GRAM) {
        for (int j = tid; j < (P / 32) * STATS_PAD; j += P) (&block_acc[0][0])[j] = 0.0;
    }
    if (tid < 20) {
        // flag expansion table: array a in {chemo app, radio app, radio dosage, death, recovery}, index = the
        // array's bit of the even column | its bit of the odd column << 1
        const int a = tid >> 2, idx = tid & 3;
        const double one = (a == 2) ? c.radio_amt : 1.0;
        lut[tid] = make_double2((idx & 1) ? one : 0.0, (idx & 2) ? one : 0.0);
    }
    __syncthreads();

    // constants -> registers
    WsK k;
    k.f = fm::consts();
    k.sphere = c.sphere; k.inv_sphere = c.inv_sphere; k.decay = c.decay; k.dose = c.chemo_amt; k.death = c.death;
    k.ndensity = -c.density; k.inv15 = fm::kInvN[15];
    if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 7; ++i) pin(k.f.lg[i]);
#pragma unroll
        for (int i = 0; i < 12; ++i) pin(k.f.ec[i]);
        pin(k.f.ln2hi); pin(k.f.ln2lo); pin(k.f.sqrt2); pin(k.f.log2e);
        pin(k.sphere); pin(k.inv_sphere); pin(k.decay); pin(k.dose); pin(k.death); pin(k.ndensity); pin(k.inv15);
    }

    const uint32_t off0 = (uint32_t)rid * 128u + (((uint32_t)rid & 7u) << 4);   // swizzled 16-byte unit 0 of the row
    uint8_t *flags_s = flag_buf + rid * Cfg::FLAG_PITCH;
    double *dummy = reinterpret_cast<double *>(dummy_buf + rid * 16);
    WsPatient p;
    WsState s;
    WsSlow slow;
    double w[18];
#pragma unroll
    for (int j = 0; j < 18; ++j) w[j] = 0.0;
    p.K = fm::log_num(1.0); p.rho = p.beta_c = p.rd = p.nb = p.si = 0.0;
    s.V = s.Cq = s.zq = s.ucp = s.udp = s.S = 0.0; s.flp = 0u; s.alive = false; s.t_end = 0;
    double v0 = 0.0;
    const int Tm1 = T - 1;
    uint32_t phase = 0;
    // fused population statistics (GRAM): per-patient sums, folded into the CTA accumulators once per tile
    PatientGram pg;
    Moments mom;
    double gVm1 = 0.0, gVm2 = 0.0;      // V[t0-1], V[t0-2] of the next group of four
    unsigned gcm2 = 0u, g_nra = 0u;     // treatment of column t0-2; radio applications so far
    int g_tlast = -4;                   // first column of the last simulated group
    const double inv_dt = 1.0 / c.fd_dt;

    for (int64_t wi = blockIdx.x; wi < n_items; wi += gridDim.x) {
        const Item it = item_of(wi);
        // ---- the item's patients: lane <-> row ----
        const int64_t patient = SKEW ? it.tile * 128 + ncls * rid + it.j : it.tile * P + tid;
        const bool exists = patient < n;
        bool tile_slow = false;
        if (MODE == 0) {
            const int64_t pi = exists ? patient : 0;
            v0 = exists ? __ldg(params + pi) : 1.0;
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
            // fast-path preconditions for this patient: one sigmoid, its argument stays inside exp_fast's
            // domain for every reachable mean diameter, K and V0 positive and normal, window of 15
            const double vmax = fmax(v0, c.death);
            const double dmax = 2.02 * cbrt(vmax * c.inv_sphere);
            const double zmax = fabs(rb) * fmax(fabs(ri), fabs(dmax - ri));
            const bool ok = (ci == ri) && (cb == rb) && (zmax <= 700.0) && (Kcap > 1e-300) && (Kcap < 1e300) &&
                            (v0 > 1e-300) && (v0 < 1e300) && (c.window == 15);
            const bool bad = exists && !ok;
            tile_slow = (P == 32) ? (__any_sync(0xffffffffu, bad) != 0) : (__syncthreads_or(bad) != 0);
            if (tile_slow) {
                slow.p = load_patient(params, pstride, pi);
                state_init(slow.s, exists);
                if (GRAM || SIDE) { slow.pg.clear(); slow.mom.clear(); }
            }
            if (GRAM || SIDE) {
                pg.clear(); mom.clear();
                gVm1 = gVm2 = 0.0; gcm2 = 0u; g_nra = 0u; g_tlast = -4;
            }
        }

        for (int ch = 0; ch < nch_all; ++ch) {
          if (ch < it.nch) {
            mbar_wait(&full_bar, phase);
            phase ^= 1u;
            const int t_first = ch * TCH - it.shift;             // global column of the chunk's first tile column
            int nb = it.nboxes - ch * NB;
            nb = nb > NB ? NB : nb;

            if (MODE == 0 && tile_slow) {
                ws_slow_chunk<P, GRAM || SIDE>(tiles, flags_s, rid, t_first, nb * 16, T, c, &slow);
            } else {
                // first body of the chunk recomputes the previous column's treatment; its outputs go to scratch
                double *prevC = dummy, *prevP = dummy + 1;
                int t_done = -1;                               // last column this chunk has simulated
                auto run_quads = [&](auto fill_tag, uint8_t *box, int tb, int lb, int h0, int h1) {
                    constexpr bool FILL = decltype(fill_tag)::value;
#pragma unroll 1
                    for (int h = h0; h < h1; ++h) {
                        const uint32_t offA = off0 ^ ((uint32_t)h << 5), offB = offA ^ 16u;
                        double2 *pa[4], *pb[4];
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            pa[a] = reinterpret_cast<double2 *>(box + a * Cfg::TILE_BYTES + offA);
                            pb[a] = reinterpret_cast<double2 *>(box + a * Cfg::TILE_BYTES + offB);
                        }
                        const double2 nzA = *pa[0], urA = *pa[1], ucA = *pa[2], udA = *pa[3];
                        const double2 nzB = *pb[0], urB = *pb[1], ucB = *pb[2], udB = *pb[3];
                        const int t0 = tb + h * 4;
                        double oV[4], oC[4], oP[4];
                        unsigned oF[4];
                        if (MODE == 1) {
                            oV[0] = nzA.x; oV[1] = nzA.y; oV[2] = nzB.x; oV[3] = nzB.y;
                            oC[0] = urA.x; oC[1] = urA.y; oC[2] = urB.x; oC[3] = urB.y;
                            oP[0] = udA.x; oP[1] = udA.y; oP[2] = udB.x; oP[3] = udB.y;
                            oF[0] = oF[1] = oF[2] = oF[3] = 0u;
                        } else {
                            ws_body<FILL, 0>(t0, Tm1, k, p, s, w, v0, nzA.x, urA.x, ucA.x, udA.x, oV[0], oC[0], oP[0], oF[0]);
                            ws_body<FILL, 1>(t0 + 1, Tm1, k, p, s, w, v0, nzA.y, urA.y, ucA.y, udA.y, oV[1], oC[1], oP[1], oF[1]);
                            ws_body<FILL, 2>(t0 + 2, Tm1, k, p, s, w, v0, nzB.x, urB.x, ucB.x, udB.x, oV[2], oC[2], oP[2], oF[2]);
                            ws_body<FILL, 3>(t0 + 3, Tm1, k, p, s, w, v0, nzB.y, urB.y, ucB.y, udB.y, oV[3], oC[3], oP[3], oF[3]);
#pragma unroll
                            for (int j = 0; j < 14; ++j) w[j] = w[j + 4];
                            if (t0 + 3 > s.t_end) {
                                // columns after the patient's last simulated one stay zero (rare: death / recovery
                                // / t = T-1).  oV[j] belongs to column t0+j, the treatment outputs to column t0+j-1.
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    if (t0 + j > s.t_end) oV[j] = 0.0;
                                    if (t0 + j - 1 > s.t_end) { oC[j] = oP[j] = 0.0; oF[j] = 0u; }
                                }
                            }
                            if (GRAM) {
                                // regression samples k = t0-2 .. t0+1 are complete now: x[k+1] and the treatment of
                                // column k+1 are known, columns after the last simulated one are already zero (so
                                // is x[seq_len], the sample the reference's snippet code appends)
                                const unsigned c0 = oF[0] & 3u, c1 = oF[1] & 3u, c2 = oF[2] & 3u, c3 = oF[3] & 3u;
                                const int te = s.t_end;
                                ws_gram_sample(pg, t0 >= 2 && t0 - 2 <= te, t0 - 2 == te || c0 != gcm2, gVm2, gVm1, gcm2,
                                               c.fd_dt, inv_dt);
                                ws_gram_sample(pg, t0 >= 1 && t0 - 1 <= te, t0 - 1 == te || c1 != c0, gVm1, oV[0], c0,
                                               c.fd_dt, inv_dt);
                                ws_gram_sample(pg, t0 <= te, t0 == te || c2 != c1, oV[0], oV[1], c1, c.fd_dt, inv_dt);
                                ws_gram_sample(pg, t0 + 1 <= te, t0 + 1 == te || c3 != c2, oV[1], oV[2], c2, c.fd_dt,
                                               inv_dt);
                                gVm2 = oV[2]; gVm1 = oV[3]; gcm2 = c3; g_tlast = t0;
                            }
                            if (GRAM || SIDE) {
                                // moments of get_scaling_params: inactive entries are zero, so plain sums do
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    mom.sv += oV[j]; mom.svv += oV[j] * oV[j];
                                    mom.sc += oC[j]; mom.scc += oC[j] * oC[j];
                                    g_nra += (oF[j] >> 1) & 1u;
                                }
                            }
                        }
                        *pa[0] = make_double2(oV[0], oV[1]);
                        *pb[0] = make_double2(oV[2], oV[3]);
                        *prevC = oC[0]; *prevP = oP[0];
                        *pa[1] = make_double2(oC[1], oC[2]);
                        *pa[3] = make_double2(oP[1], oP[2]);
                        pb[1]->x = oC[3];
                        pb[3]->x = oP[3];
                        prevC = &pb[1]->y; prevP = &pb[3]->y;
                        uint8_t *fd = flags_s + 3 + lb + h * 4;    // byte of column t0 - 1 (byte 3 = dummy)
                        fd[0] = (uint8_t)oF[0]; fd[1] = (uint8_t)oF[1]; fd[2] = (uint8_t)oF[2]; fd[3] = (uint8_t)oF[3];
                        t_done = t0 + 3;
                    }
                };
                for (int b = 0; b < nb; ++b) {
                    uint8_t *box = tiles + b * 4 * Cfg::TILE_BYTES;
                    const int tb = t_first + b * 16;
                    // quads left of column 0 (skewed first box) and right of the horizon are skipped; the window
                    // fills during columns 0..15, which have their own code copy
                    int h_lo = tb < 0 ? (-tb) >> 2 : 0;
                    int h_hi = (T - tb + 3) >> 2;
                    h_hi = h_hi > 4 ? 4 : h_hi;
                    int h_fill = (16 - tb) >> 2;                 // quads with t0 < 16
                    h_fill = h_fill < h_lo ? h_lo : (h_fill > h_hi ? h_hi : h_fill);
                    if (h_fill > h_lo) run_quads(std::true_type{}, box, tb, b * 16, h_lo, h_fill);
                    if (h_hi > h_fill) run_quads(std::false_type{}, box, tb, b * 16, h_fill, h_hi);
                }
                // treatment of the chunk's last simulated column (the next chunk's first body repeats it from the
                // same state)
                if (t_done >= 0) {
                    double pr = 0.0, C1 = 0.0;
                    unsigned f = 0u;
                    if (MODE == 0 && t_done <= s.t_end && t_done >= 1) {
                        bool ra, ca;
                        ws_treat(k, s, pr, ra, ca, C1);
                        f = s.flp | (ca ? 1u : 0u) | (ra ? 2u : 0u);
                    }
                    *prevC = C1; *prevP = pr;
                    flags_s[4 + (t_done - t_first)] = (uint8_t)f;
                }
            }
            fence_proxy_async_smem();
            cta_sync();
            const int src_pc = (MODE == 0 && tile_slow) ? 2 : 3;   // fast path: one sigmoid serves both arrays
            if (SKEW && ch == 0 && it.shift > 0) {
                // head of the rows: the 16 - shift columns before the first line boundary, tile columns
                // [shift, 16) of box 0, by hand (16-byte stores; lanes <-> column pairs, rows in turn)
                const int npairs = (16 - it.shift) >> 1;
                const int64_t row0 = it.tile * 128 + it.j;
                for (int idx = rid; idx < 32 * npairs; idx += 32) {
                    const int row = idx / npairs, cpair = idx - row * npairs;
                    const uint32_t so = swz_off<128>((uint32_t)row, (uint32_t)((it.shift >> 1) + cpair));
                    const int64_t go = (row0 + ncls * row) * T + 2 * cpair;
                    const double2 vV = *reinterpret_cast<const double2 *>(tiles + 0 * Cfg::TILE_BYTES + so);
                    const double2 vC = *reinterpret_cast<const double2 *>(tiles + 1 * Cfg::TILE_BYTES + so);
                    const double2 vc = *reinterpret_cast<const double2 *>(tiles + src_pc * Cfg::TILE_BYTES + so);
                    const double2 vr = *reinterpret_cast<const double2 *>(tiles + 3 * Cfg::TILE_BYTES + so);
                    __stcs(reinterpret_cast<double2 *>(out_V + go), vV);
                    __stcs(reinterpret_cast<double2 *>(out_C + go), vC);
                    __stcs(reinterpret_cast<double2 *>(out_pc + go), vc);
                    __stcs(reinterpret_cast<double2 *>(out_pr + go), vr);
                }
                __syncwarp();
            }
            if (leader) {
                const uint64_t pol = opt_evict_first ? l2_policy_evict_first() : 0ull;
                for (int b = 0; b < nb; ++b) {
                    uint8_t *box = tiles + b * 4 * Cfg::TILE_BYTES;
                    const int m = ch * NB + b;
                    if (SKEW && it.shift > 0 && m == 0) continue;   // stored by hand above
                    if (16 * m - it.shift < T) {
                        store_box(box + 0 * Cfg::TILE_BYTES, 0, it, m, pol);
                        store_box(box + 1 * Cfg::TILE_BYTES, 1, it, m, pol);
                        store_box(box + src_pc * Cfg::TILE_BYTES, 2, it, m, pol);
                        store_box(box + 3 * Cfg::TILE_BYTES, 3, it, m, pol);
                    }
                }
                tma_store_commit();
                tma_store_wait_read();   // the tiles are free again
            }
          }
            // the next chunk's loads go out before the flag expansion, which then overlaps their latency; the
            // row-class warps issue them together (see above)
            if (SKEW) __syncthreads();
            if (leader) {
                if (ch + 1 < it.nch) issue_load(it, ch + 1);
                else if (ch + 1 == nch_all && wi + gridDim.x < n_items) issue_load(item_of(wi + gridDim.x), 0);
            }
          if (ch < it.nch) {
            const int t_first = ch * TCH - it.shift;
            int nb = it.nboxes - ch * NB;
            nb = nb > NB ? NB : nb;
            if (SIDE && exists) {
                // treatment codes chemo + 2*radio of the chunk's columns: the thread's own flag bytes, 16 per store
                // (code_pitch is a multiple of 16 that covers T rounded up, so whole boxes fit)
                const uint32_t *fw = reinterpret_cast<const uint32_t *>(flags_s + 4);
                uint8_t *crow = codes_out + patient * code_pitch + t_first;
                for (int b = 0; b < nb; ++b) {
                    uint4 v;
                    v.x = fw[4 * b + 0] & 0x03030303u; v.y = fw[4 * b + 1] & 0x03030303u;
                    v.z = fw[4 * b + 2] & 0x03030303u; v.w = fw[4 * b + 3] & 0x03030303u;
                    *reinterpret_cast<uint4 *>(crow + 16 * b) = v;
                }
            }
            // expand the packed flag bytes: item = (row, column pair); consecutive lanes write consecutive 16 bytes.
            // Each array's double2 comes from a 4-entry table indexed by its bit of the two columns.
            {
                static_assert(P >= HALF, "flag expansion mapping: every thread keeps one column pair");
                constexpr int RSTEP = P / HALF;                      // rows advanced per iteration
                const int64_t row0 = SKEW ? it.tile * 128 + it.j : it.tile * P;
                const int rmul = SKEW ? ncls : 1;                        // global rows per local row
                const int rows_valid = SKEW ? 32 : ((n - row0 < P) ? (int)(n - row0) : P);
                const int cp = rid % HALF, r_first = rid / HALF;
                const int col = t_first + cp * 2;
                if (col >= 0 && col < T && cp * 2 < nb * 16) {   // T, shift even: a column pair is in or out as a whole
                    const int64_t tile_ofs = row0 * pitch + col;
                    double *b_ca = out_ca + tile_ofs, *b_ra = out_ra + tile_ofs, *b_D = out_D + tile_ofs,
                           *b_de = out_death + tile_ofs, *b_re = out_recov + tile_ofs;
                    const uint8_t *fp = flag_buf + r_first * Cfg::FLAG_PITCH + cp * 2 + 4;
                    int ofs = r_first * rmul * (int)pitch;
#pragma unroll 4
                    for (int row = r_first; row < rows_valid;
                         row += RSTEP, ofs += RSTEP * rmul * (int)pitch, fp += RSTEP * Cfg::FLAG_PITCH) {
                        const unsigned two = *reinterpret_cast<const unsigned short *>(fp);
                        // bits: 1 chemo, 2 radio, 4 death, 8 recovery; low byte = even column, high byte = odd column
                        const unsigned i_ca = (two & 1u) | ((two >> 7) & 2u);
                        const unsigned i_ra = ((two >> 1) & 1u) | ((two >> 8) & 2u);
                        const unsigned i_de = ((two >> 2) & 1u) | ((two >> 9) & 2u);
                        const unsigned i_re = ((two >> 3) & 1u) | ((two >> 10) & 2u);
                        const double2 v_ca = lut[0 + i_ca], v_ra = lut[4 + i_ra], v_D = lut[8 + i_ra],
                                      v_de = lut[12 + i_de], v_re = lut[16 + i_re];
                        if (opt_evict_first) {
                            __stcs(reinterpret_cast<double2 *>(b_ca + ofs), v_ca);
                            __stcs(reinterpret_cast<double2 *>(b_ra + ofs), v_ra);
                            __stcs(reinterpret_cast<double2 *>(b_D + ofs), v_D);
                            __stcs(reinterpret_cast<double2 *>(b_de + ofs), v_de);
                            __stcs(reinterpret_cast<double2 *>(b_re + ofs), v_re);
                        } else {
                            *reinterpret_cast<double2 *>(b_ca + ofs) = v_ca;
                            *reinterpret_cast<double2 *>(b_ra + ofs) = v_ra;
                            *reinterpret_cast<double2 *>(b_D + ofs) = v_D;
                            *reinterpret_cast<double2 *>(b_de + ofs) = v_de;
                            *reinterpret_cast<double2 *>(b_re + ofs) = v_re;
                        }
                    }
                }
            }
            cta_sync();   // every thread is done with the flag bytes before the next chunk rewrites them
          }
        }
        if (MODE == 0 && exists) seq_len_out[patient] = (double)((tile_slow ? slow.s.t_end : s.t_end) + 1);
        if (SIDE && exists) {
            const Moments &mm = tile_slow ? slow.mom : mom;
            const double sd = tile_slow ? slow.mom.sd : c.radio_amt * (double)g_nra;
            const double sdd = tile_slow ? slow.mom.sdd : c.radio_amt * c.radio_amt * (double)g_nra;
            pmom_out[0 * pstride + patient] = mm.sv;  pmom_out[1 * pstride + patient] = mm.svv;
            pmom_out[2 * pstride + patient] = mm.sc;  pmom_out[3 * pstride + patient] = mm.scc;
            pmom_out[4 * pstride + patient] = sd;     pmom_out[5 * pstride + patient] = sdd;
        }
        if (GRAM) {
            const double u = exists ? __ldg(static_feature + patient) : 0.0;
            const int lane = tid & 31, wrp = tid >> 5;
            if (tile_slow) {
                factual_finish<true>(c, slow.s, slow.pg);
                fold_patient_stats(block_acc[wrp], lane, slow.pg, slow.mom, u, exists, slow.s.t_end + 1);
            } else {
                // the last sample (k = seq_len - 1) when it lies in the second half of the last group
                ws_gram_sample(pg, g_tlast + 2 <= s.t_end, true, gVm2, gVm1, gcm2, c.fd_dt, inv_dt);
                mom.sd = c.radio_amt * (double)g_nra;
                mom.sdd = c.radio_amt * c.radio_amt * (double)g_nra;
                fold_patient_stats(block_acc[wrp], lane, pg, mom, u, exists, s.t_end + 1);
            }
        }
    }
    if (leader) tma_store_wait_all();
    if (GRAM) stats_block_finish(block_acc, P / 32, ws, &s_is_last);
}

// tuning switches of the data-movement skeleton; B200I_WS_OPTS overrides the default
static int ws_env_opts()
{
    const char *e = getenv("B200I_WS_OPTS");
    return e ? atoi(e) : 6;   // measured best: evict-first on output stores, evict-last on draw loads
}

static int ws_encode(TmapPack &pack, int64_t n, int T, int64_t pitch, int P, const double *const in[4],
                     double *const out[9])
{
    for (int a = 0; a < 4; ++a) {
        int rc = encode_tmap_2d_pitched_f64(&pack.in[a], in[a], (uint64_t)n, (uint64_t)T, (uint64_t)pitch * 8, P, 16);
        if (rc) return rc;
    }
    for (int a = 0; a < 9; ++a) {
        int rc = encode_tmap_2d_pitched_f64(&pack.out[a], out[a], (uint64_t)n, (uint64_t)T, (uint64_t)pitch * 8, P, 16);
        if (rc) return rc;
    }
    return 0;
}
// per-class views of the (n, T) arrays, n a multiple of 4: class j = rows j, j+4, ... with a pitch of 4 rows
static int ws_encode(TmapPackSkew &pack, int64_t n, int T, int64_t pitch, int ncls, const double *const in[4],
                     double *const out[9])
{
    B200I_REQUIRE(pitch == T, B200I_E_UNSUPPORTED, "row-class mapping: dense rows only (pitch %lld, T %d)", (long long)pitch, T);
    const int oidx[4] = {0, 1, 5, 6};   // volume, chemo dosage, chemo / radio probabilities
    for (int j = 0; j < ncls; ++j) {
        const int shift = (j * T) & (4 * ncls - 1), head = (16 - shift) & 15;
        for (int a = 0; a < 4; ++a) {
            int rc = encode_tmap_2d_pitched_f64(&pack.in[j][a], in[a] + (j * T - shift), (uint64_t)(n / ncls),
                                                (uint64_t)(T + shift), (uint64_t)T * 8 * ncls, 32, 16);
            if (rc) return rc;
            rc = encode_tmap_2d_pitched_f64(&pack.out[j][a], out[oidx[a]] + (j * T + head), (uint64_t)(n / ncls),
                                            (uint64_t)(T - head), (uint64_t)T * 8 * ncls, 32, 16);
            if (rc) return rc;
        }
    }
    return 0;
}

template <int P, int NB, int MINB, int MODE, bool SKEW, int STATS = 0>
static int launch_ws(int64_t n, int64_t pstride, int T, int64_t pitch, const SimC &c, const double *params,
                     const double *const in[4], double *const out[9], double *seq_len, cudaStream_t st,
                     const double *static_feature = nullptr, StatsWorkspace *ws = nullptr, uint8_t *codes_out = nullptr,
                     int64_t code_pitch = 0, double *pmom_out = nullptr)
{
    using Cfg = WsCfg<P, NB>;
    if (n <= 0) return 0;
    typename WsMaps<SKEW>::type pack;
    {
        int rc = ws_encode(pack, n, T, pitch, SKEW ? 4 : P, in, out);
        if (rc) return rc;
    }
    constexpr bool GRAM = STATS == 1;
    auto kern = sim_factual_ws<P, NB, MINB, MODE, SKEW, STATS>;
    constexpr int SMEM = SKEW ? Cfg::SMEM_BYTES_4 : Cfg::SMEM_BYTES;
    constexpr int NT = SKEW ? 128 : P;
    B200I_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    int per_sm = 0;
    B200I_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, SMEM));
    B200I_REQUIRE(per_sm >= 1, B200I_E_UNSUPPORTED, "sim_factual_ws<%d,%d>: does not fit on an SM", P, NB);
    const int64_t n_items = SKEW ? n / 128 : (n + P - 1) / P;
    int64_t grid = (int64_t)num_sms() * per_sm;
    if (grid > n_items) grid = n_items;
    if (GRAM && grid > STATS_MAX_BLOCKS) grid = STATS_MAX_BLOCKS;
    static const int opts = ws_env_opts();
    // out order: V C D ca ra pc pr death recov
    kern<<<(unsigned)grid, NT, SMEM, st>>>(pack, opts, n, pstride, T, pitch, c, params, out[3], out[4], out[2], out[7],
                                                     out[8], seq_len, out[0], out[1], out[5], out[6], static_feature, ws, codes_out,
                                                     code_pitch, pmom_out);
    return check_cuda(cudaGetLastError(), "sim_factual_ws launch");
}

}  // namespace b200i
