// theta_gram.cu -- K4: the data reduction behind SINDY.fit for the cancer simulator.
//
// Reference semantics: process_sindy_training_data (pkpd/utils.py:433-462) cuts
// every patient's trajectory x[0..L] (L = sequence_length, x[L] included) into maximal constant-
// treatment snippets that share their end point with the next snippet; pysindy differentiates each
// snippet with FiniteDifference(order=1) (forward difference, last point backward) and evaluates
// PolynomialLibrary(degree=2, interaction_only=True) = [1, x0, u0, x0*u0]; the four per-treatment
// regressions (sindy.py:203-213) only need Theta^T Theta and Theta^T xdot.  This kernel streams the
// cohort once and reduces those normal equations in FP64; nothing tall is ever materialised.
//
// Mapping: CTA of 128 threads = 128 consecutive patients.  Their volume rows are one contiguous
// chunk of 128*T doubles, fetched with a single 1-D bulk TMA copy (cp.async.bulk, SASS UBLKCP) into
// shared memory; the two application arrays are read with coalesced 16-byte loads and packed to one
// treatment-code byte per step.  Then one thread walks one patient.  HBM traffic = 3 arrays once.
// The moments needed by get_scaling_params (cancer_simulation.py:776-796) come from a second,
// element-parallel kernel over the same arrays (masked_moments).
#include "fastmath.cuh"
#include "sim_math.cuh"
#include "stats_reduce.cuh"
#include "tma.cuh"
#include <stdlib.h>

namespace b200i {

constexpr int GP = 128;  // patients per CTA

__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <bool BULK>
__global__ void __launch_bounds__(GP)
theta_gram_kernel(int64_t n, int T, double fd_dt, const double *__restrict__ vol, const double *__restrict__ chemo,
                  const double *__restrict__ radio, const double *__restrict__ seq_len,
                  const double *__restrict__ static_feature, StatsWorkspace *ws, const double *__restrict__ dts = nullptr,
                  int dts_per_row = 0)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ double block_acc[STATS_MAX_WARPS][STATS_PAD];
    __shared__ unsigned int s_is_last;
    double *s_vol = reinterpret_cast<double *>(smem_raw);                         // [GP][T]
    uint8_t *s_code = smem_raw + (size_t)GP * T * sizeof(double);                  // [GP][T]
    // irregular sampling: interval lengths dts[k] = t[k+1] - t[k], (T-1,) for the cohort or (n, T-1) per patient
    double *s_dt = reinterpret_cast<double *>(smem_raw + (((size_t)GP * T * 9 + 15) & ~(size_t)15));   // [GP][T-1] or [T-1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (dts && !dts_per_row)
        for (int k = tid; k < T - 1; k += GP) s_dt[k] = dts[k];

    for (int j = tid; j < STATS_MAX_WARPS * STATS_PAD; j += GP) (&block_acc[0][0])[j] = 0.0;
    if (BULK && tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t ntiles = (n + GP - 1) / GP;
    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t first = tile * GP;
        const int rows = (int)((n - first < GP) ? (n - first) : GP);
        const int64_t elems = (int64_t)rows * T;
        const double *gv = vol + first * T;
        if (BULK) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&bar, (uint32_t)(elems * 8));
                bulk_load_1d(s_vol, gv, (uint32_t)(elems * 8), &bar);
            }
        } else {
            for (int64_t e = tid; e < elems; e += GP) s_vol[e] = gv[e];
        }
        if (dts && dts_per_row) {
            const double *gd = dts + first * (T - 1);
            for (int64_t e = tid; e < (int64_t)rows * (T - 1); e += GP) s_dt[e] = gd[e];
        }
        // treatment codes: code = chemo + 2*radio (dataset.py:130-141 one-hot index)
        const double *gc = chemo + first * T, *gr = radio + first * T;
        if (BULK) {  // T even, 16-byte aligned: two steps per load
            const double2 *gc2 = reinterpret_cast<const double2 *>(gc), *gr2 = reinterpret_cast<const double2 *>(gr);
            for (int64_t e = tid; e < elems / 2; e += GP) {
                const double2 a = __ldg(gc2 + e), b = __ldg(gr2 + e);
                s_code[2 * e] = (uint8_t)((a.x != 0.0 ? 1 : 0) + (b.x != 0.0 ? 2 : 0));
                s_code[2 * e + 1] = (uint8_t)((a.y != 0.0 ? 1 : 0) + (b.y != 0.0 ? 2 : 0));
            }
        } else {
            for (int64_t e = tid; e < elems; e += GP)
                s_code[e] = (uint8_t)((gc[e] != 0.0 ? 1 : 0) + (gr[e] != 0.0 ? 2 : 0));
        }
        __syncthreads();
        if (BULK) {
            mbar_wait(&bar, phase);
            phase ^= 1u;
        }
        const bool exists = tid < rows;
        PatientGram pg;
        pg.clear();
        int L = 0;
        double u = 0.0;
        if (exists) {
            L = (int)seq_len[first + tid];
            if (L > T - 1) L = T - 1;
            u = static_feature[first + tid];
            const double *x = s_vol + (size_t)tid * T;
            const uint8_t *a = s_code + (size_t)tid * T;
            const double *dtr = dts ? s_dt + (dts_per_row ? (size_t)tid * (T - 1) : 0) : nullptr;
            double x0 = x[0];
            int a0 = a[0];
            for (int k = 0; k < L; ++k) {
                const double x1 = x[k + 1];
                const int a1 = a[k + 1];
                const double xdot = __ddiv_rn(__dsub_rn(x1, x0), dtr ? dtr[k] : fd_dt);
                pg.add(a0, x0, xdot);
                if (k == L - 1 || a1 != a0) pg.add(a0, x1, xdot);
                x0 = x1;
                a0 = a1;
            }
        }
        // fold (Gram part only; moments are produced by masked_moments_kernel)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            double g[B200I_GRAM_PER_TREATMENT];
            expand_gram(pg.s[a], u, g);
#pragma unroll
            for (int j = 0; j < B200I_GRAM_PER_TREATMENT; ++j)
                warp_acc_add(block_acc[warp], a * B200I_GRAM_PER_TREATMENT + j, exists ? g[j] : 0.0, lane);
        }
        __syncthreads();  // smem tile is reused by the next iteration
    }
    stats_block_finish(block_acc, GP >> 5, ws, &s_is_last);
}

// ------------------------------------------------------------------------------------------------
// theta_gram, second generation (default when T is even and the arrays are 16-byte aligned): ONE pass over
// the five arrays produces both the normal equations and the scaling moments.
//
// The first generation (theta_gram_kernel + masked_moments_kernel) took 0.73 + 0.35 ms at 1M patients
// (profiles/r1_launches_first.csv): two launches that both read the volume array, a thread walking a
// 480-byte-pitch row in shared memory (8-way bank conflicts) and a serial load -> pack -> walk per CTA.
//   * warp = 32 consecutive patients; every warp of the CTA runs its own tile pipeline (own mbarrier)
//   * volume rows arrive by per-row bulk copies (cp.async.bulk, one row per lane) into a 16-byte padded
//     pitch, so the thread-per-patient walk reads two columns per conflict-free LDS.128
//   * while they are in flight the warp streams the four other arrays with coalesced 16-byte loads:
//     applications -> one treatment-code byte per step (transposed [T][33] layout: conflict-free both
//     ways), dosages -> masked moments on the fly
//   * the per-patient sums are expanded with the static feature and kept in registers across tiles; the
//     shuffle / shared-memory / ordered-grid reduction runs once per thread at the end
// ------------------------------------------------------------------------------------------------
constexpr int G2_WARPS = 4;
// packed slot -> (power of the static feature u, per-patient sum {n, sum x, sum x^2, sum xdot, sum x xdot}); cf. expand_gram
__device__ const unsigned char kSlotPower[B200I_GRAM_PER_TREATMENT] = {0, 0, 1, 1, 0, 1, 1, 2, 2, 2, 0, 0, 1, 1, 0};
__device__ const unsigned char kSlotMoment[B200I_GRAM_PER_TREATMENT] = {0, 1, 0, 1, 2, 1, 2, 0, 1, 2, 3, 4, 3, 4, 0};

// CODES: the lean fit of the device pipeline -- treatment codes (one byte per step) and six per-patient moment sums
// written by the simulator kernel (b200i_sim_factual_side) replace the two application and two dosage arrays:
// 0.6 instead of 2.4 GB per million patients.
template <int MAXT, bool CODES>
__global__ void __launch_bounds__(G2_WARPS * 32, CODES ? 3 : 2)   // the five-array form keeps 32 loads in flight per thread
theta_gram2_kernel(int64_t n, int T, int64_t rp, int mode, double fd_dt, double inv_dt, const double *__restrict__ vol,
                   const double *__restrict__ chemo, const double *__restrict__ radio,
                   const double *__restrict__ seq_len, const double *__restrict__ static_feature,
                   const double *__restrict__ chemo_dos, const double *__restrict__ radio_dos, StatsWorkspace *ws,
                   const uint8_t *__restrict__ codes, int64_t code_pitch, const double *__restrict__ pmom, int64_t mstride)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint64_t bars[G2_WARPS];
    __shared__ double block_acc[G2_WARPS][STATS_PAD];
    __shared__ unsigned int s_is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pitch = T * 8 + 16;                                   // volume row pitch (bytes)
    const int warp_bytes = 32 * pitch + ((T * 33 + 15) & ~15);
    uint8_t *s_vol = smem_raw + (size_t)warp * warp_bytes;          // [32][pitch]
    uint8_t *s_code = s_vol + 32 * pitch;                           // [T][33]

    for (int j = tid; j < G2_WARPS * STATS_PAD; j += blockDim.x) (&block_acc[0][0])[j] = 0.0;
    if (lane == 0) mbar_init(&bars[warp], 32);
    mbar_fence_init();
    __syncthreads();

    // Population sums.  A patient contributes its 20 sums s[a][m] times 1, u, u^2 (u = static feature).  Instead of
    // 60 expanded accumulators per thread (120 registers, which capped the kernel at 8 warps per SM), the 32
    // patients of a tile are reduced through shared memory: lane j < 20 owns sum j = (treatment, moment) and adds
    // the tile's 32 values in patient order with the three weights -- 3 accumulators per lane, fixed order.
    double A0 = 0.0, A1 = 0.0, A2 = 0.0;
    double mv = 0, mvv = 0, mc = 0, mcc = 0, md = 0, mdd = 0, mcnt = 0, mrows = 0;
    double *s_red = reinterpret_cast<double *>(s_vol);   // [20][33] sums + [32] u: aliases the volume tile after the walk

    const int64_t ntiles = (n + 31) / 32;
    const int half = T / 2;
    uint32_t phase = 0;
    for (int64_t tile = (int64_t)blockIdx.x * G2_WARPS + warp; tile < ntiles; tile += (int64_t)gridDim.x * G2_WARPS) {
        const int64_t first = tile * 32;
        const int rows = (int)((n - first < 32) ? (n - first) : 32);
        // volume rows: lane j copies row j
        if (lane < rows) {
            mbar_arrive_expect_tx(&bars[warp], (uint32_t)(T * 8));
            bulk_load_1d(s_vol + lane * pitch, vol + (first + lane) * rp, (uint32_t)(T * 8), &bars[warp]);
        } else {
            mbar_arrive(&bars[warp]);
        }
        if (CODES) {
            // everything this tile needs from global memory is requested before the first dependent instruction:
            // the per-patient scalars, then the code bytes in batches of four 16-byte loads per lane (item = (row,
            // 16-byte unit)), which are transposed into the [T][33] table while the volume rows are still in flight
            const int64_t pidx = first + (lane < rows ? lane : 0);
            const double p0 = __ldg(pmom + 0 * mstride + pidx), p1 = __ldg(pmom + 1 * mstride + pidx);
            const double p2 = __ldg(pmom + 2 * mstride + pidx), p3 = __ldg(pmom + 3 * mstride + pidx);
            const double p4 = __ldg(pmom + 4 * mstride + pidx), p5 = __ldg(pmom + 5 * mstride + pidx);
            const int units = (T + 15) >> 4, items = rows * units;
            for (int e0 = 0; e0 < items; e0 += 128) {
                uint4 v[4];
                int jj[4], uu[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int e = e0 + 32 * b + lane;
                    const int ec = e < items ? e : 0;
                    jj[b] = ec / units; uu[b] = ec - jj[b] * units;
                    v[b] = __ldg(reinterpret_cast<const uint4 *>(codes + (first + jj[b]) * code_pitch) + uu[b]);
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (e0 + 32 * b + lane < items) {
                        const unsigned wv[4] = {v[b].x, v[b].y, v[b].z, v[b].w};
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const int col = uu[b] * 16 + q;
                            if (col < T) s_code[col * 33 + jj[b]] = (uint8_t)((wv[q >> 2] >> (8 * (q & 3))) & 3u);
                        }
                    }
                }
            }
            if (lane < rows) {   // the simulator's per-patient sums over the active entries
                mv += p0; mvv += p1; mc += p2; mcc += p3; md += p4; mdd += p5;
            }
        }
        // the four other arrays, two columns per lane and row
        const double2 *gc2 = reinterpret_cast<const double2 *>(chemo + first * rp);
        const double2 *gr2 = reinterpret_cast<const double2 *>(radio + first * rp);
        const double2 *gC2 = chemo_dos ? reinterpret_cast<const double2 *>(chemo_dos + first * rp) : nullptr;
        const double2 *gD2 = radio_dos ? reinterpret_cast<const double2 *>(radio_dos + first * rp) : nullptr;
        const int64_t rp2 = rp / 2;
        int L_lane = (lane < rows) ? (int)__ldg(seq_len + first + lane) : 0;
        L_lane = L_lane > T ? T : L_lane;
        for (int cb = 0; !CODES && cb < half; cb += 32) {
            const bool active = cb + lane < half;          // every lane iterates (warp shuffles below)
            const int c = active ? cb + lane : 0;
            // batches of 8 rows: all 32 loads of a batch are issued before the first dependent store
            for (int j0 = 0; j0 < rows; j0 += 8) {
                double2 va[8], vb[8], vC[8], vD[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = (j0 + i < rows) ? j0 + i : rows - 1;
                    const int64_t e = (int64_t)j * rp2 + c;
                    va[i] = __ldg(gc2 + e); vb[i] = __ldg(gr2 + e);
                    vC[i] = gC2 ? __ldg(gC2 + e) : make_double2(0.0, 0.0);
                    vD[i] = gD2 ? __ldg(gD2 + e) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = j0 + i;
                    const int L = __shfl_sync(0xffffffffu, L_lane, j & 31);
                    if (active && j < rows) {
                        s_code[(2 * c) * 33 + j] = (uint8_t)((va[i].x != 0.0 ? 1 : 0) + (vb[i].x != 0.0 ? 2 : 0));
                        s_code[(2 * c + 1) * 33 + j] = (uint8_t)((va[i].y != 0.0 ? 1 : 0) + (vb[i].y != 0.0 ? 2 : 0));
                        const double w0 = (2 * c < L) ? 1.0 : 0.0, w1 = (2 * c + 1 < L) ? 1.0 : 0.0;
                        mc += w0 * vC[i].x + w1 * vC[i].y;
                        mcc += w0 * vC[i].x * vC[i].x + w1 * vC[i].y * vC[i].y;
                        md += w0 * vD[i].x + w1 * vD[i].y;
                        mdd += w0 * vD[i].x * vD[i].x + w1 * vD[i].y * vD[i].y;
                    }
                }
            }
        }
        __syncwarp();
        mbar_wait(&bars[warp], phase);
        phase ^= 1u;
        // thread-per-patient walk over the snippets (pkpd/utils.py:433-462 + FiniteDifference order 1)
        if (lane < rows) {
            int Ls = (int)seq_len[first + lane];
            Ls = Ls > T ? T : Ls;
            const int L = Ls > T - 1 ? T - 1 : Ls;
            const double u = static_feature[first + lane];
            const uint8_t *row = s_vol + lane * pitch;
            PatientGram pg;
            pg.clear();
            // one regression sample (x0 -> x1 under treatment a) plus, when the snippet ends at x1, its
            // backward-difference end point: both rows share xdot and the treatment, so they are merged and
            // filed with one-hot weights -- branch-free, 20 independent FMA chains
            auto sample = [&](double x0, double x1, int a, bool end) {
                const double xdot = fm::div_small(__dsub_rn(x1, x0), fd_dt, inv_dt);
                const double e = end ? 1.0 : 0.0;
                const double cnt = 1.0 + e, sx = fma(e, x1, x0), sxx = fma(e * x1, x1, x0 * x0);
                const double sd = cnt * xdot, sxd = sx * xdot;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const double wt = (a == t) ? 1.0 : 0.0;
                    pg.s[t][0] = fma(wt, cnt, pg.s[t][0]); pg.s[t][1] = fma(wt, sx, pg.s[t][1]);
                    pg.s[t][2] = fma(wt, sxx, pg.s[t][2]); pg.s[t][3] = fma(wt, sd, pg.s[t][3]);
                    pg.s[t][4] = fma(wt, sxd, pg.s[t][4]);
                }
            };
            if (mode == 1) {
                // joint model (process_sindy_training_data(joint=True), pkpd/utils.py:493-497 with the arguments of
                // :656-672): ONE trajectory per patient, x_k = V[k+1] (the outputs), inputs of sample k = the
                // applications of column k; FiniteDifference order 1 over the whole trajectory (last point backward)
                const double *xr = reinterpret_cast<const double *>(row);
                double xd_prev = 0.0;
                for (int k = 0; k < L; ++k) {
                    const double x = xr[k + 1];
                    const double xdot = (k < L - 1) ? fm::div_small(__dsub_rn(xr[k + 2], x), fd_dt, inv_dt) : xd_prev;
                    const int a = s_code[k * 33 + lane];
                    const double sxx = x * x, sxd = x * xdot;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const double wt = (a == t) ? 1.0 : 0.0;
                        pg.s[t][0] += wt; pg.s[t][1] = fma(wt, x, pg.s[t][1]);
                        pg.s[t][2] = fma(wt, sxx, pg.s[t][2]); pg.s[t][3] = fma(wt, xdot, pg.s[t][3]);
                        pg.s[t][4] = fma(wt, sxd, pg.s[t][4]);
                    }
                    xd_prev = xdot;
                }
                if (!CODES) for (int k = 0; k < Ls; ++k) { mv += xr[k]; mvv += xr[k] * xr[k]; }
            }
            double2 cur = *reinterpret_cast<const double2 *>(row);
            int a0 = s_code[lane];
            for (int k = 0; mode == 0 && k < L; k += 2) {
                const double2 nxt = *reinterpret_cast<const double2 *>(row + (k + 2) * 8);   // pitch padding keeps it in bounds
                const int a1 = s_code[(k + 1) * 33 + lane];
                const int a2 = (k + 2 < T) ? s_code[(k + 2) * 33 + lane] : 0;
                sample(cur.x, cur.y, a0, k == L - 1 || a1 != a0);
                if (!CODES) { mv += cur.x; mvv += cur.x * cur.x; }
                if (k + 1 < L) {
                    sample(cur.y, nxt.x, a1, k + 1 == L - 1 || a2 != a1);
                    if (!CODES) { mv += cur.y; mvv += cur.y * cur.y; }
                }
                a0 = a2;
                cur = nxt;
            }
            if (!CODES && mode == 0 && Ls > L) {   // sequence_length == T: the last entry is active but starts no sample
                const double x = *reinterpret_cast<const double *>(row + (size_t)L * 8);
                mv += x; mvv += x * x;
            }
            mcnt += (double)Ls; mrows += 1.0;
            __syncwarp();   // every lane has finished reading its volume row
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int m = 0; m < 5; ++m) s_red[(a * 5 + m) * 33 + lane] = pg.s[a][m];
            s_red[20 * 33 + lane] = u;
        } else {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 20; ++j) s_red[j * 33 + lane] = 0.0;
            s_red[20 * 33 + lane] = 0.0;
        }
        __syncwarp();
        if (lane < 20) {
            const double *mine = s_red + lane * 33, *us = s_red + 20 * 33;
#pragma unroll 8
            for (int q = 0; q < 32; ++q) {
                const double v = mine[q], uq = us[q];
                A0 += v;
                A1 = fma(uq, v, A1);
                A2 = fma(uq * uq, v, A2);
            }
        }
        __syncwarp();   // the warp's tile is reused by its next iteration
    }
    // lane j holds the sums of (treatment j / 5, moment j % 5) with weights 1, u, u^2 -> packed layout of b200i.h:
    // G00 G01 G02 G03 G11 G12 G13 G22 G23 G33 | b0 b1 b2 b3 | count  per treatment (Theta = [1, x, u, x u])
    {
        double *fin = s_red;   // [3][20]
        __syncwarp();
        if (lane < 20) { fin[lane] = A0; fin[20 + lane] = A1; fin[40 + lane] = A2; }
        __syncwarp();
        for (int idx = lane; idx < 4 * B200I_GRAM_PER_TREATMENT; idx += 32) {
            const int a = idx / B200I_GRAM_PER_TREATMENT, sl = idx - a * B200I_GRAM_PER_TREATMENT;
            block_acc[warp][idx] += fin[kSlotPower[sl] * 20 + a * 5 + kSlotMoment[sl]];
        }
        __syncwarp();
    }
    const int m0 = 4 * B200I_GRAM_PER_TREATMENT;
    warp_acc_add(block_acc[warp], m0 + 0, mv, lane);  warp_acc_add(block_acc[warp], m0 + 1, mvv, lane);
    warp_acc_add(block_acc[warp], m0 + 2, mc, lane);  warp_acc_add(block_acc[warp], m0 + 3, mcc, lane);
    warp_acc_add(block_acc[warp], m0 + 4, md, lane);  warp_acc_add(block_acc[warp], m0 + 5, mdd, lane);
    warp_acc_add(block_acc[warp], m0 + 6, mcnt, lane); warp_acc_add(block_acc[warp], m0 + 7, mrows, lane);
    stats_block_finish(block_acc, G2_WARPS, ws, &s_is_last);
}

// sum x, sum x^2 over active entries [i, :seq_len[i]] of up to three (N,T) arrays; one warp per row.
// Writes moments into a second StatsWorkspace-style partial area (slots 60..67 of the same layout).
__global__ void __launch_bounds__(256)
masked_moments_kernel(int64_t n, int T, const double *__restrict__ a0, const double *__restrict__ a1,
                      const double *__restrict__ a2, const double *__restrict__ seq_len, StatsWorkspace *ws)
{
    __shared__ double block_acc[STATS_MAX_WARPS][STATS_PAD];
    __shared__ unsigned int s_is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int j = tid; j < STATS_MAX_WARPS * STATS_PAD; j += blockDim.x) (&block_acc[0][0])[j] = 0.0;
    __syncthreads();
    double s[6] = {0, 0, 0, 0, 0, 0};
    double cnt = 0.0, rows = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * nwarps + warp; i < n; i += (int64_t)gridDim.x * nwarps) {
        int L = (int)seq_len[i];
        if (L > T) L = T;
        for (int k = lane; k < L; k += 32) {
            const double v = a0[i * T + k];
            s[0] += v; s[1] += v * v;
            if (a1) { const double c = a1[i * T + k]; s[2] += c; s[3] += c * c; }
            if (a2) { const double d = a2[i * T + k]; s[4] += d; s[5] += d * d; }
        }
        if (lane == 0) { cnt += (double)L; rows += 1.0; }
    }
    const int m0 = 4 * B200I_GRAM_PER_TREATMENT;
#pragma unroll
    for (int j = 0; j < 6; ++j) warp_acc_add(block_acc[warp], m0 + j, s[j], lane);
    warp_acc_add(block_acc[warp], m0 + 6, cnt, lane);
    warp_acc_add(block_acc[warp], m0 + 7, rows, lane);
    // reduce only the moment slots; Gram slots of this launch are zero and must not clobber stats[0..59]
    __syncthreads();
    if (tid < B200I_MOMENTS) {
        double v = 0.0;
        for (int w = 0; w < nwarps; ++w) v += block_acc[w][m0 + tid];
        ws->partials[blockIdx.x][m0 + tid] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(&ws->ticket[1], 1u);
        s_is_last = (t == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_is_last) {
        __threadfence();
        if (tid < B200I_MOMENTS) {
            double v = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) v += __ldcg(&ws->partials[b][m0 + tid]);
            ws->stats[m0 + tid] = v;
        }
        if (tid == 0) ws->ticket[1] = 0u;
    }
}

}  // namespace b200i

using namespace b200i;

extern "C" int b200i_theta_gram(int64_t n, int32_t T, double fd_dt, const double *cancer_volume,
                                const double *chemo_application, const double *radio_application,
                                const double *sequence_lengths, const double *static_feature,
                                const double *chemo_dosage, const double *radio_dosage, void *gram_workspace,
                                void *stream)
{
    return b200i_theta_gram_pitched(n, T, T, fd_dt, cancer_volume, chemo_application, radio_application, sequence_lengths,
                                    static_feature, chemo_dosage, radio_dosage, gram_workspace, stream);
}

extern "C" int b200i_theta_gram_pitched(int64_t n, int32_t T, int64_t row_pitch, double fd_dt,
                                        const double *cancer_volume, const double *chemo_application,
                                        const double *radio_application, const double *sequence_lengths,
                                        const double *static_feature, const double *chemo_dosage,
                                        const double *radio_dosage, void *gram_workspace, void *stream)
{
    return b200i_theta_gram_mode(n, T, row_pitch, 0, fd_dt, cancer_volume, chemo_application, radio_application,
                                 sequence_lengths, static_feature, chemo_dosage, radio_dosage, gram_workspace, stream);
}

extern "C" int b200i_theta_gram_codes(int64_t n, int32_t T, int64_t row_pitch, int32_t mode, double fd_dt,
                                      const double *cancer_volume, const uint8_t *codes, int64_t code_pitch,
                                      const double *sequence_lengths, const double *static_feature,
                                      const double *patient_moments, int64_t moments_stride, void *gram_workspace,
                                      void *stream)
{
    if (moments_stride == 0) moments_stride = n;
    B200I_REQUIRE(moments_stride >= n, B200I_E_ARG, "theta_gram_codes: moments_stride %lld < n", (long long)moments_stride);
    B200I_REQUIRE(n >= 0 && gram_workspace, B200I_E_ARG, "theta_gram_codes: negative n or NULL workspace");
    if (n == 0) {   // an empty cohort (its arrays may be NULL): zero statistics
        B200I_CUDA(cudaMemsetAsync(gram_workspace, 0, sizeof(double) * 128 + sizeof(unsigned int) * 32,
                                   static_cast<cudaStream_t>(stream)));
        return 0;
    }
    B200I_REQUIRE(cancer_volume && codes && sequence_lengths && static_feature && patient_moments, B200I_E_ARG,
                  "theta_gram_codes: NULL argument");
    B200I_REQUIRE(T >= 2 && T <= 256 && T % 2 == 0 && fd_dt > 0 && (mode == 0 || mode == 1), B200I_E_UNSUPPORTED,
                  "theta_gram_codes: T=%d (even, <= 256), fd_dt > 0, mode 0/1", T);
    B200I_REQUIRE(row_pitch >= T && (row_pitch == T || row_pitch % 2 == 0) && code_pitch >= ((T + 15) / 16) * 16 &&
                      code_pitch % 16 == 0 && aligned16(cancer_volume) && aligned16(codes),
                  B200I_E_ARG, "theta_gram_codes: row_pitch %lld / code_pitch %lld (multiple of 16, >= T rounded up)",
                  (long long)row_pitch, (long long)code_pitch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StatsWorkspace *ws = static_cast<StatsWorkspace *>(gram_workspace);
    B200I_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 128 + sizeof(unsigned int) * 32, st));
    if (n == 0) return 0;
    const size_t warp_bytes = (size_t)32 * (T * 8 + 16) + (((size_t)T * 33 + 15) & ~(size_t)15);
    const size_t smem2 = warp_bytes * G2_WARPS;
    auto k2 = theta_gram2_kernel<256, true>;
    int per_sm2 = 0;   // attribute + occupancy cached process-wide: this launch is issued per chunk by the upload pipeline
    {
        int rc = ensure_dyn_smem(reinterpret_cast<const void *>(k2), (int)smem2, G2_WARPS * 32, &per_sm2);
        if (rc) return rc;
    }
    B200I_REQUIRE(per_sm2 >= 1, B200I_E_UNSUPPORTED, "theta_gram_codes: T=%d does not fit in shared memory", T);
    const int64_t ntiles2 = (n + 31) / 32;
    int64_t grid2 = (int64_t)num_sms() * per_sm2;
    const int64_t need = (ntiles2 + G2_WARPS - 1) / G2_WARPS;
    if (grid2 > need) grid2 = need;
    if (grid2 > STATS_MAX_BLOCKS) grid2 = STATS_MAX_BLOCKS;
    k2<<<(unsigned)grid2, G2_WARPS * 32, smem2, st>>>(n, T, row_pitch, mode, fd_dt, 1.0 / fd_dt, cancer_volume, nullptr,
                                                      nullptr, sequence_lengths, static_feature, nullptr, nullptr, ws,
                                                      codes, code_pitch, patient_moments, moments_stride);
    return check_cuda(cudaGetLastError(), "theta_gram_codes launch");
}

extern "C" int b200i_theta_gram_mode(int64_t n, int32_t T, int64_t row_pitch, int32_t mode, double fd_dt,
                                     const double *cancer_volume, const double *chemo_application,
                                     const double *radio_application, const double *sequence_lengths,
                                     const double *static_feature, const double *chemo_dosage,
                                     const double *radio_dosage, void *gram_workspace, void *stream)
{
    B200I_REQUIRE(mode == 0 || mode == 1, B200I_E_ARG, "theta_gram: mode %d (0 = per-treatment snippets, 1 = joint)", mode);
    B200I_REQUIRE(row_pitch >= T && (row_pitch == T || row_pitch % 2 == 0), B200I_E_ARG,
                  "theta_gram: row_pitch %lld (T = %d) must be even and >= T", (long long)row_pitch, T);
    B200I_REQUIRE(n >= 0 && cancer_volume && chemo_application && radio_application && sequence_lengths &&
                      static_feature && gram_workspace,
                  B200I_E_ARG, "theta_gram: NULL argument or negative n");
    B200I_REQUIRE(T >= 2 && T <= 1024, B200I_E_UNSUPPORTED, "theta_gram: T=%d outside [2,1024]", T);
    B200I_REQUIRE(fd_dt > 0, B200I_E_ARG, "theta_gram: fd_dt must be positive");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StatsWorkspace *ws = static_cast<StatsWorkspace *>(gram_workspace);
    // stats + tickets start from zero for every call (partials are fully overwritten)
    B200I_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 128 + sizeof(unsigned int) * 32, st));
    if (n == 0) return 0;
    {
        const bool v2 = (T % 2 == 0) && T <= 256 && aligned16(cancer_volume) && aligned16(chemo_application) &&
                        aligned16(radio_application) && (!chemo_dosage || aligned16(chemo_dosage)) &&
                        (!radio_dosage || aligned16(radio_dosage)) &&
                        (row_pitch != T || mode != 0 || getenv("B200I_THETA_GRAM_V1") == nullptr);
        if (v2) {
            const size_t warp_bytes = (size_t)32 * (T * 8 + 16) + (((size_t)T * 33 + 15) & ~(size_t)15);
            const size_t smem2 = warp_bytes * G2_WARPS;
            auto k2 = theta_gram2_kernel<256, false>;
            B200I_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            int per_sm2 = 0;
            B200I_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k2, G2_WARPS * 32, smem2));
            if (per_sm2 >= 1) {
                const int64_t ntiles2 = (n + 31) / 32;
                int64_t grid2 = (int64_t)num_sms() * per_sm2;
                const int64_t need = (ntiles2 + G2_WARPS - 1) / G2_WARPS;
                if (grid2 > need) grid2 = need;
                if (grid2 > STATS_MAX_BLOCKS) grid2 = STATS_MAX_BLOCKS;
                k2<<<(unsigned)grid2, G2_WARPS * 32, smem2, st>>>(n, T, row_pitch, mode, fd_dt, 1.0 / fd_dt, cancer_volume, chemo_application,
                                                                  radio_application, sequence_lengths, static_feature,
                                                                  chemo_dosage, radio_dosage, ws, nullptr, 0, nullptr, 0);
                return check_cuda(cudaGetLastError(), "theta_gram2 launch");
            }
        }
    }
    B200I_REQUIRE(row_pitch == T && mode == 0, B200I_E_UNSUPPORTED,
                  "theta_gram: pitched rows and the joint mode need even T <= 256 and 16-byte aligned arrays");
    const size_t smem = (size_t)GP * T * 9;
    const bool bulk = (T % 2 == 0) && aligned16(cancer_volume) && aligned16(chemo_application) &&
                      aligned16(radio_application);
    auto kern = bulk ? theta_gram_kernel<true> : theta_gram_kernel<false>;
    B200I_REQUIRE(smem <= 220 * 1024, B200I_E_UNSUPPORTED, "theta_gram: T=%d needs %zu bytes of shared memory", T, smem);
    B200I_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    B200I_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GP, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t ntiles = (n + GP - 1) / GP;
    int64_t grid = (int64_t)num_sms() * per_sm;
    if (grid > ntiles) grid = ntiles;
    if (grid > STATS_MAX_BLOCKS) grid = STATS_MAX_BLOCKS;
    kern<<<(unsigned)grid, GP, smem, st>>>(n, T, fd_dt, cancer_volume, chemo_application, radio_application,
                                           sequence_lengths, static_feature, ws, nullptr, 0);
    B200I_CUDA(cudaGetLastError());
    int64_t g2 = (n + 7) / 8;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (g2 > cap) g2 = cap;
    masked_moments_kernel<<<(unsigned)g2, 256, 0, st>>>(n, T, cancer_volume, chemo_dosage, radio_dosage,
                                                        sequence_lengths, ws);
    return check_cuda(cudaGetLastError(), "theta_gram launch");
}

// Population statistics on an irregular time grid (BASELINE config C4): finite differences
// (x[k+1] - x[k]) / (t[k+1] - t[k]), pysindy FiniteDifference(order=1) with a time array instead of a scalar step
// (the reference always passes the uniform STANDARD_DT, sindy.py:195,203-213; odeint itself takes any grid,
// pkpd/utils.py:68-90).  dts = interval lengths, (T-1,) for the cohort or (n, T-1) per patient.  Gram part only
// (no dosage moments); first-generation kernel.
extern "C" int b200i_theta_gram_dts(int64_t n, int32_t T, const double *cancer_volume, const double *chemo_application,
                                    const double *radio_application, const double *sequence_lengths,
                                    const double *static_feature, const double *dts, int32_t dts_per_row,
                                    void *gram_workspace, void *stream)
{
    B200I_REQUIRE(n >= 0 && cancer_volume && chemo_application && radio_application && sequence_lengths && static_feature &&
                      dts && gram_workspace,
                  B200I_E_ARG, "theta_gram_dts: NULL argument or negative n");
    B200I_REQUIRE(T >= 2 && T <= 1024, B200I_E_UNSUPPORTED, "theta_gram_dts: T=%d outside [2,1024]", T);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StatsWorkspace *ws = static_cast<StatsWorkspace *>(gram_workspace);
    B200I_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 128 + sizeof(unsigned int) * 32, st));
    if (n == 0) return 0;
    const size_t smem = (((size_t)GP * T * 9 + 15) & ~(size_t)15) + (size_t)(dts_per_row ? GP : 1) * (T - 1) * sizeof(double);
    B200I_REQUIRE(smem <= 220 * 1024, B200I_E_UNSUPPORTED, "theta_gram_dts: T=%d needs %zu bytes of shared memory", T, smem);
    auto kern = theta_gram_kernel<false>;
    int per_sm = 1;
    {
        int rc0 = ensure_dyn_smem(reinterpret_cast<const void *>(kern), (int)smem, GP, &per_sm);
        if (rc0) return rc0;
    }
    if (per_sm < 1) per_sm = 1;
    const int64_t ntiles = (n + GP - 1) / GP;
    int64_t grid = (int64_t)num_sms() * per_sm;
    if (grid > ntiles) grid = ntiles;
    if (grid > STATS_MAX_BLOCKS) grid = STATS_MAX_BLOCKS;
    kern<<<(unsigned)grid, GP, smem, st>>>(n, T, 1.0, cancer_volume, chemo_application, radio_application, sequence_lengths,
                                           static_feature, ws, dts, dts_per_row);
    return check_cuda(cudaGetLastError(), "theta_gram_dts launch");
}

// ------------------------------------------------------------------------------------------------
// SmoothedFiniteDifference(smoother_kws={'window_length': 2, 'polyorder': 1}) pre-pass (sindy.py:196-198).
// scipy.signal.savgol_filter with an even window fits its line at the half-sample position: every interior sample of a
// trajectory becomes (x[i] + x[i+1]) / 2, the first and last sample are refitted through their two neighbours and stay
// x[i] (mode='interp').  The trajectories are the ones theta_gram cuts: per-treatment snippets share their end sample
// with the next snippet's first sample, and both are edge samples, so one smoothed copy of the row serves every snippet;
// joint model: one trajectory per patient over columns 1..L.  The smoothed row then goes through the ordinary
// statistics kernels (pysindy >= 1.7.4 builds the library from the smoothed samples as well).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) smooth_snippets_kernel(int64_t n, int T, const double *__restrict__ vol,
                                                              const double *__restrict__ chemo,
                                                              const double *__restrict__ radio,
                                                              const double *__restrict__ seq, int joint,
                                                              double *__restrict__ out)
{
    const int64_t total = n * (int64_t)T;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = idx / T;
        const int i = (int)(idx - p * T);
        const int L = (int)seq[p];
        const double x = vol[idx];
        bool edge;
        if (joint)
            edge = i <= 1 || i >= L;
        else
            edge = i == 0 || i >= L || chemo[idx] != chemo[idx - 1] || radio[idx] != radio[idx - 1];
        out[idx] = (edge || i + 1 >= T) ? x : 0.5 * x + 0.5 * vol[idx + 1];
    }
}

extern "C" int b200i_smooth_snippets(int64_t n, int32_t T, const double *cancer_volume, const double *chemo_application,
                                     const double *radio_application, const double *sequence_lengths, int32_t joint,
                                     double *smoothed_out, void *stream)
{
    B200I_REQUIRE(n >= 0, B200I_E_ARG, "smooth_snippets: negative n");
    if (n == 0) return 0;     // empty cohort: nothing to write (the arrays of an empty cohort may be NULL)
    B200I_REQUIRE(cancer_volume && chemo_application && radio_application && sequence_lengths && smoothed_out,
                  B200I_E_ARG, "smooth_snippets: NULL argument");
    B200I_REQUIRE(T >= 2, B200I_E_UNSUPPORTED, "smooth_snippets: T=%d < 2", T);
    B200I_REQUIRE(smoothed_out != cancer_volume, B200I_E_ARG, "smooth_snippets: the pass is not in-place");
    const int64_t total = n * (int64_t)T;
    int64_t grid = (total + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (grid > cap) grid = cap;
    smooth_snippets_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n, T, cancer_volume, chemo_application, radio_application, sequence_lengths, joint, smoothed_out);
    return check_cuda(cudaGetLastError(), "smooth_snippets launch");
}
