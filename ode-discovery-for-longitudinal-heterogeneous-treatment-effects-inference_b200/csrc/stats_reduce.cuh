// stats_reduce.cuh -- deterministic reduction of the packed population statistics
// (B200I_STATS_DOUBLES = 4 x 15 Gram/RHS/count + 8 moments) from threads to one vector.
//
// Workspace layout (b200i_gram_workspace_bytes()):
//   double   stats[128]                 final result (first B200I_STATS_DOUBLES used)
//   uint32_t ticket[32]                 arrival counter of the "last block reduces" pattern
//   double   partials[MAX_BLOCKS][STATS_PAD]
// Thread -> warp (shuffle tree) -> block (warp partials in shared memory, summed in warp order)
// -> grid (block partials in global memory, summed in block order by the last block to arrive).
// No floating-point atomics, so the result is bit-reproducible for a fixed launch shape.
#pragma once
#include "common.cuh"

namespace b200i {

constexpr int STATS = B200I_STATS_DOUBLES;  // 68
constexpr int STATS_PAD = 72;
constexpr int STATS_MAX_BLOCKS = 4096;
constexpr int STATS_MAX_WARPS = 8;

struct StatsWorkspace {
    double stats[128];
    unsigned int ticket[32];
    double partials[STATS_MAX_BLOCKS][STATS_PAD];
};

// adds `v` (already reduced over the warp, valid in lane 0) into the warp's accumulator slot
__device__ __forceinline__ void warp_acc_add(double *warp_acc, int slot, double v, int lane)
{
    v = warp_sum(v);
    if (lane == 0) warp_acc[slot] += v;
}

// block_acc: shared double[STATS_MAX_WARPS][STATS_PAD] holding per-warp accumulators.
// Call with all threads of the block after the last accumulation.
__device__ __forceinline__ void stats_block_finish(double (*block_acc)[STATS_PAD], int nwarps, StatsWorkspace *ws,
                                                   unsigned int *s_is_last)
{
    __syncthreads();
    const int tid = threadIdx.x;
    for (int i = tid; i < STATS; i += blockDim.x) {   // blocks smaller than STATS threads loop
        double v = 0.0;
        for (int w = 0; w < nwarps; ++w) v += block_acc[w][i];
        ws->partials[blockIdx.x][i] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(&ws->ticket[0], 1u);
        *s_is_last = (t == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (*s_is_last) {
        __threadfence();
        for (int i = tid; i < STATS; i += blockDim.x) {
            double v = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) v += __ldcg(&ws->partials[b][i]);
            ws->stats[i] = v;
        }
        if (tid == 0) ws->ticket[0] = 0u;  // ready for the next launch
    }
}

}  // namespace b200i
