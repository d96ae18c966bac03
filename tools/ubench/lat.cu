// Microbenchmark (single warp): dependent-chain latencies of DFMA / DMUL / SHFL / DMMA on sm_100a and the cost of one
// in-register 8x8 complex Gauss-Jordan inversion (gj_invert8) — the serial part of the factorisation's macro-step.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../hmcmt2d_b200/csrc/band_factor.cuh"
using namespace hmcmt;

__global__ void k_lat(double* out, long long* cyc, int iters) {
    double x = threadIdx.x * 1e-3 + 1.0, a = 1.0000001, b = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    cyc[0] = t1 - t0;
    double y = x;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) y = y * a;
    t1 = clock64();
    cyc[1] = t1 - t0;
    double z = y;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 1) & 31);
    t1 = clock64();
    cyc[2] = t1 - t0;
    double c[2] = {z, x};
    t0 = clock64();
    for (int i = 0; i < iters; ++i) dmma884(c, a, b);
    t1 = clock64();
    cyc[3] = t1 - t0;
    // the inversion: matrix = diagonally dominant complex symmetric
    const int lane = threadIdx.x & 31, i = lane >> 2, t = lane & 3;
    cplx a0 = mk((i == 2 * t) ? 4.0 : 0.1 * (i + 2 * t), 0.3), a1 = mk((i == 2 * t + 1) ? 4.0 : 0.1 * (i + 2 * t + 1), 0.2);
    bool bad = false;
    t0 = clock64();
    for (int r = 0; r < 16; ++r) { gj_invert8<true>(a0, a1, bad, i, t); a0.x += 1.0; }
    t1 = clock64();
    cyc[4] = (t1 - t0) / 16;
    out[threadIdx.x] = c[0] + c[1] + a0.x + a1.y + (bad ? 1 : 0);
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 8 * 8);
    const int iters = 4096;
    k_lat<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    k_lat<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("dependent DFMA %.1f cycles, DMUL %.1f, SHFL(double = 2 SHFL) %.1f, DMMA.8x8x4 %.1f, gj_invert8 %lld cycles\n",
           (double)cyc[0] / iters, (double)cyc[1] / iters, (double)cyc[2] / iters, (double)cyc[3] / iters, cyc[4]);
    return 0;
}
