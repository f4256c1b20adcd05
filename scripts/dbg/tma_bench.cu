// Micro-benchmark: what does the TMA path sustain for {TC columns x P rows} boxes over (N,60) f64 arrays?
// Each CTA streams tiles: 4 loads + NST stores per chunk (copying input tiles to output arrays), no compute.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include "../../ode-discovery-for-longitudinal-heterogeneous-treatment-effects-inference_b200/csrc/tma.cuh"
using namespace b200i;
namespace b200i { void set_error(const char*, ...) {} int check_cuda(cudaError_t e, const char*) { return (int)e; } int num_sms(){return 148;} }
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static enc_fn g_enc;
static void mk(CUtensorMap* m, void* base, uint64_t n, uint64_t T, uint32_t P, uint32_t TC) {
    cuuint64_t gd[2] = {T, n}; cuuint64_t gs[1] = {T*8}; cuuint32_t box[2] = {TC, P}; cuuint32_t es[2] = {1,1};
    CUtensorMapSwizzle sw = TC*8==32?CU_TENSOR_MAP_SWIZZLE_32B: TC*8==64?CU_TENSOR_MAP_SWIZZLE_64B: TC*8==128?CU_TENSOR_MAP_SWIZZLE_128B: CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode failed %d\n", (int)r); exit(1); }
}
struct Maps { CUtensorMap in[4]; CUtensorMap out[9]; };

template <int NBUF>
__global__ void k(const __grid_constant__ Maps maps, long n, int T, int P, int TC, int NST, int delay) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bars[NBUF];
    uint8_t* tiles = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    const int tile_bytes = P * TC * 8;
    const int nchunks = (T + TC - 1) / TC;
    const long ntiles = (n + P - 1) / P;
    const long my = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long total = my * nchunks;
    if (threadIdx.x == 0) { for (int b=0;b<NBUF;++b) mbar_init(&bars[b], 1); mbar_fence_init(); }
    __syncthreads();
    auto load = [&](long g) {
        int b = g % NBUF; long tile = blockIdx.x + (g / nchunks) * gridDim.x; int ch = g % nchunks;
        mbar_arrive_expect_tx(&bars[b], 4u * tile_bytes);
        for (int a=0;a<4;++a) tma_load_2d(tiles + (b*4+a)*tile_bytes, &maps.in[a], ch*TC, (int)(tile*P), &bars[b]);
    };
    if (threadIdx.x == 0) for (long g=0; g<NBUF && g<total; ++g) load(g);
    for (long g = 0; g < total; ++g) {
        int b = g % NBUF; long tile = blockIdx.x + (g / nchunks) * gridDim.x; int ch = g % nchunks;
        mbar_wait(&bars[b], (uint32_t)((g / NBUF) & 1));
        // fake compute
        long long t0 = clock64(); while (clock64() - t0 < delay) {}
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int a=0;a<NST;++a) tma_store_2d(&maps.out[a], ch*TC, (int)(tile*P), tiles + (b*4 + (a&3))*tile_bytes);
            tma_store_commit();
            tma_store_wait_read();
            if (g + NBUF < total) load(g + NBUF);
        }
    }
    if (threadIdx.x == 0) tma_store_wait_all();
}

int main(int argc, char** argv) {
    long n = 1000000; int T = 60;
    void* p=nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q); g_enc = (enc_fn)p;
    double *in[4], *out[9];
    for (int a=0;a<4;++a) { cudaMalloc(&in[a], n*T*8); cudaMemset(in[a], 0, n*T*8); }
    for (int a=0;a<9;++a) cudaMalloc(&out[a], n*T*8);
    int cfgs[][5] = { // P, TC, NST, NBUF, ctas/SM
        {128,8,9,1,2},{128,8,4,1,4},{128,8,4,1,2},{128,8,4,2,2},{128,16,4,1,2},{128,16,4,2,1},{64,16,4,1,4},{64,16,4,2,2},
        {32,16,4,1,8},{128,4,4,1,4},{128,4,4,1,8},{32,60,4,1,3},{32,60,9,1,1},{64,30,4,1,2},{128,8,0,1,4},{128,8,0,2,4},{128,16,0,1,2},{128,16,0,2,2}};
    for (auto& c : cfgs) {
        int P=c[0], TC=c[1], NST=c[2], NBUF=c[3], per=c[4];
        Maps m;
        for (int a=0;a<4;++a) mk(&m.in[a], in[a], n, T, P, TC);
        for (int a=0;a<9;++a) mk(&m.out[a], out[a], n, T, P, TC);
        int smem = NBUF*4*P*TC*8 + 1024;
        auto kern = NBUF==1 ? k<1> : k<2>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int delay : {0, 20000}) {
            cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            kern<<<148*per, 128, smem>>>(m, n, T, P, TC, NST, delay); cudaDeviceSynchronize();
            cudaEventRecord(e0);
            kern<<<148*per, 128, smem>>>(m, n, T, P, TC, NST, delay);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            cudaError_t err = cudaGetLastError();
            double gb = (4.0 + NST) * n * T * 8 / 1e9;
            printf("P=%3d TC=%2d NST=%d NBUF=%d ctas/SM=%d smem=%6d delay=%5d : %.3f ms  %.0f GB/s %s\n", P, TC, NST, NBUF, per, smem, delay, ms, gb/ms*1e3, err?cudaGetErrorString(err):"");
        }
    }
    return 0;
}
