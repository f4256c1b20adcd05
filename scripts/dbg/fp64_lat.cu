// Micro-benchmark: dependent-issue latency and per-SM throughput of the FP64 pipe, MUFU.RCP64H, I2F/F2F on B200.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void lat(double *out, long long *cyc, double a, double b, int iters)
{
    double x = a + threadIdx.x * 1e-9, y = b;
    double x2 = x + 1.0, x3 = x + 2.0, x4 = x + 3.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (MODE == 0) x = fma(x, y, a);                 // dependent DFMA
            if (MODE == 1) x = __dadd_rn(x, y);              // dependent DADD
            if (MODE == 2) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
            if (MODE == 3) x = (double)(float)x;             // F2F down + up
            if (MODE == 4) x = (double)(__double2int_rn(x)); // D2I + I2F
            if (MODE == 5) { x = fma(x, y, a); x2 = fma(x2, y, a); }                    // 2 chains
            if (MODE == 6) { x = fma(x, y, a); x2 = fma(x2, y, a); x3 = fma(x3, y, a); x4 = fma(x4, y, a); }  // 4 chains
            if (MODE == 7) x = (double)exp2f(__log2f((float)x) * 0.3333f);   // cbrt seed chain
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + x2 + x3 + x4;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    const char *names[] = {"DFMA dep", "DADD dep", "MUFU.RCP64H dep", "F2F down+up dep", "D2I+I2F dep", "DFMA 2 chains", "DFMA 4 chains", "cbrt seed chain"};
    const int ops[] = {16, 16, 16, 16, 16, 32, 64, 16};
    for (int warps : {1, 4, 8, 16}) {
        printf("-- %d warps per SM (1 block)\n", warps);
        for (int m = 0; m < 8; ++m) {
            int iters = 256;
            void (*k)(double *, long long *, double, double, int) = nullptr;
            switch (m) { case 0: k = lat<0>; break; case 1: k = lat<1>; break; case 2: k = lat<2>; break; case 3: k = lat<3>; break;
                         case 4: k = lat<4>; break; case 5: k = lat<5>; break; case 6: k = lat<6>; break; case 7: k = lat<7>; break; }
            k<<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999, iters);
            k<<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999, iters);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-18s %7.2f cycles per op-slot (%d ops per slot-group) -> %.2f warp-ops/cycle/SM\n", names[m], (double)h / (iters * 16), ops[m] / 16,
                   (double)warps * iters * ops[m] / h);
        }
    }
    return 0;
}
