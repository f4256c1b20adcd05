// common.cuh -- shared host/device helpers of libb200insite (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/b200i.h"

namespace b200i {

void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
int num_sms();
// Process-wide, thread-safe launch configuration of a kernel with dynamic shared memory: raises the kernel's
// MaxDynamicSharedMemorySize to `bytes` if it is lower (never lowers it, so concurrent callers with different sizes
// cannot invalidate each other's launches) and returns the resident CTAs per SM for (threads, bytes) in *per_sm.
int ensure_dyn_smem(const void *func, int bytes, int threads, int *per_sm);
// Stream-ordered scratch memory from a library-owned pool (one per device) whose release threshold keeps freed blocks
// cached: the default pool hands its memory back to the driver at every synchronisation, which turns a 480 MB
// scratch buffer into milliseconds of cudaMalloc per call.
int pool_alloc(void **ptr, size_t bytes, cudaStream_t stream);
void pool_free(void *ptr, cudaStream_t stream);

#define B200I_CUDA(call)                                   \
    do {                                                   \
        int _rc = ::b200i::check_cuda((call), #call);      \
        if (_rc) return _rc;                               \
    } while (0)

#define B200I_REQUIRE(cond, code, ...)                     \
    do {                                                   \
        if (!(cond)) {                                     \
            ::b200i::set_error(__VA_ARGS__);               \
            return (code);                                 \
        }                                                  \
    } while (0)

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- warp / block reductions (FP64, fixed order => deterministic) ------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace b200i
