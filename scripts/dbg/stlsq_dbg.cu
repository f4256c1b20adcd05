#include <cstdio>
#include "../../ode-discovery-for-longitudinal-heterogeneous-treatment-effects-inference_b200/csrc/stlsq.cuh"
using namespace b200i;
__global__ void k(const double* g15, double thr, double alpha) {
    double G[4][4], b[4], c[4];
    unpack_gram(g15, G, b);
    printf("G row0 %g %g %g %g  b %g %g %g %g\n", G[0][0], G[0][1], G[0][2], G[0][3], b[0], b[1], b[2], b[3]);
    bool ok = solve_spd4(G, b, 0xF, alpha, nullptr, c);
    printf("ridge ok=%d c %g %g %g %g\n", (int)ok, c[0], c[1], c[2], c[3]);
    ok = solve_spd4(G, b, 0xF, 0.0, nullptr, c);
    printf("ols ok=%d c %g %g %g %g\n", (int)ok, c[0], c[1], c[2], c[3]);
    unsigned ind = stlsq4(G, b, thr, alpha, 100, 0xFu, c);
    printf("stlsq ind=%u c %g %g %g %g\n", ind, c[0], c[1], c[2], c[3]);
}
int main() {
    double h[15] = {4.2092e4, 3.396606939885e+05, 8.3487e4, 6.745574146596e+05, 1.572528690466e+08, 6.745574146596e+05, 3.111246504730e+08, 1.93947e5, 1.551384475495e+06, 7.084356640584e+08, 1.022107428780e+03, 4.481985352219e+06, 1.530820877116e+04, 1.565461677065e+07, 42092};
    double* d; cudaMalloc(&d, sizeof h); cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
    k<<<1,1>>>(d, 1e-3, 0.5); cudaDeviceSynchronize();
    double G[4][4], b[4], c[4];
    unpack_gram(h, G, b);
    unsigned ind = stlsq4(G, b, 1e-3, 0.5, 100, 0xFu, c);
    printf("host ind=%u c %g %g %g %g\n", ind, c[0], c[1], c[2], c[3]);
    return 0;
}
