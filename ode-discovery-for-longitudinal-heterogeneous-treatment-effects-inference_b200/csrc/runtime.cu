// runtime.cu -- error reporting, device queries and tensor-map encoding shared by all kernels.
#include <cuda.h>
#include <stdarg.h>
#include <string.h>
#include <map>
#include <mutex>
#include <tuple>
#include "common.cuh"
#include "tma.cuh"

namespace b200i {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return static_cast<int>(e);
}

int num_sms()
{
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
        cached = prop.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

int ensure_dyn_smem(const void *func, int bytes, int threads, int *per_sm)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, int> raised;                   // (kernel, device) -> attribute value
    static std::map<std::tuple<const void *, int, int, int>, int> occupancy;     // (kernel, device, threads, bytes)
    int dev = 0;
    B200I_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    int &cur = raised[{func, dev}];
    if (bytes > cur) {
        B200I_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        cur = bytes;
    }
    if (per_sm) {
        auto key = std::make_tuple(func, dev, threads, bytes);
        auto it = occupancy.find(key);
        if (it == occupancy.end()) {
            int v = 0;
            B200I_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, func, threads, (size_t)bytes));
            it = occupancy.emplace(key, v).first;
        }
        *per_sm = it->second;
    }
    return 0;
}

static cudaMemPool_t g_pools[64];
static std::mutex g_pool_mu;

int pool_alloc(void **ptr, size_t bytes, cudaStream_t stream)
{
    int dev = 0;
    B200I_CUDA(cudaGetDevice(&dev));
    B200I_REQUIRE(dev >= 0 && dev < 64, B200I_E_UNSUPPORTED, "pool_alloc: device index %d", dev);
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        if (g_pools[dev] == nullptr) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            B200I_CUDA(cudaMemPoolCreate(&g_pools[dev], &props));
            uint64_t keep = UINT64_MAX;
            B200I_CUDA(cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
        }
        pool = g_pools[dev];
    }
    B200I_CUDA(cudaMallocFromPoolAsync(ptr, bytes, pool, stream));
    return 0;
}

void pool_free(void *ptr, cudaStream_t stream)
{
    if (ptr) cudaFreeAsync(ptr, stream);
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode()
{
    // function-local static: initialised exactly once, thread-safe (C++11 magic statics)
    static const encode_tiled_fn fn = []() -> encode_tiled_fn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return reinterpret_cast<encode_tiled_fn>(p);
        return nullptr;
    }();
    return fn;
}

int encode_tmap_2d_f64(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                       uint32_t box_cols, bool promote_256)
{
    encode_tiled_fn enc = get_encode();
    B200I_REQUIRE(enc != nullptr, B200I_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
    const uint32_t row_bytes = box_cols * 8u;
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
    if (row_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
    else if (row_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
    else if (row_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
    else if (row_bytes % 16 != 0 || box_cols > 256) {
        set_error("tensor map: box of %u columns (rows of %u bytes) is not supported", box_cols, row_bytes);
        return B200I_E_UNSUPPORTED;
    }   // any other multiple of 16 bytes: unswizzled rows
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 8u};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     promote_256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu box=%ux%u base=%p)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, box_rows, box_cols, base);
        return B200I_E_DRIVER;
    }
    return 0;
}

int encode_tmap_2d_pitched_f64(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                               uint32_t box_rows, uint32_t box_cols)
{
    encode_tiled_fn enc = get_encode();
    B200I_REQUIRE(enc != nullptr, B200I_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
    B200I_REQUIRE(box_cols * 8u == 128u && pitch_bytes % 16 == 0 && rows > 0 && cols > 0 && aligned16(base),
                  B200I_E_UNSUPPORTED, "tensor map (pitched): box of %u columns, pitch %llu, base %p", box_cols,
                  (unsigned long long)pitch_bytes, base);
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (pitched) failed with CUresult %d (rows=%llu cols=%llu pitch=%llu base=%p)",
                  (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_bytes, base);
        return B200I_E_DRIVER;
    }
    return 0;
}

int encode_tmap_3d_rowgroups_f64(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint32_t group_rows,
                                 uint32_t box_groups, uint32_t box_cols)
{
    encode_tiled_fn enc = get_encode();
    B200I_REQUIRE(enc != nullptr, B200I_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
    B200I_REQUIRE(box_cols * 8u == 128u && rows % group_rows == 0 && rows > 0, B200I_E_UNSUPPORTED,
                  "tensor map (row groups): box of %u columns, %llu rows in groups of %u", box_cols,
                  (unsigned long long)rows, group_rows);
    cuuint64_t gdim[3] = {cols, group_rows, rows / group_rows};
    cuuint64_t gstride[2] = {cols * 8u, cols * 8u * group_rows};
    cuuint32_t box[3] = {box_cols, 1, box_groups};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult %d (rows=%llu cols=%llu base=%p)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, base);
        return B200I_E_DRIVER;
    }
    return 0;
}

}  // namespace b200i

extern "C" const char *b200i_last_error(void) { return b200i::g_err; }
extern "C" int b200i_version(void) { return 100; }
extern "C" int b200i_device_sms(int *sms_out)
{
    if (!sms_out) return B200I_E_ARG;
    *sms_out = b200i::num_sms();
    return 0;
}
