// Microbenchmark: FP64 DFMA vs DMMA (m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16) throughput on sm_100a,
// plus LDS.128 broadcast cost.  Scratch tool used to choose the factorisation kernel design.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

__global__ void k_dfma(double* out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void k_dmma884(double* out, int iters) {
    double c[NT][2];
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#ifdef BIGMMA
template <int NT>
__global__ void k_dmma1688(double* out, int iters) {
    double c[NT][4];
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 1.0 + threadIdx.x * 1e-4, b1 = b0 * 2;
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 1; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

// LDS.128 cost: mode 0 = all lanes same address, 1 = 8 distinct addresses (lane/4), 2 = 32 distinct consecutive
__global__ void k_lds(double* out, int iters, int mode) {
    extern __shared__ double2 sm[];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_double2(i, -i);
    __syncthreads();
    int lane = threadIdx.x & 31;
    int base = mode == 0 ? 0 : (mode == 1 ? (lane >> 2) : lane);
    double2 acc = make_double2(0, 0);
    int idx = base;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double2 v = sm[(idx + u * 64) & 2047];
            acc.x += v.x; acc.y += v.y;
        }
        idx = (idx + (int)acc.x * 0 + 32) & 2047;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int threads : {128, 256, 512, 1024}) {
        int iters = 20000;
        k_dfma<<<sms, threads>>>(out, 100, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); k_dfma<<<sms, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 16 * iters * (double)threads * sms;
        printf("DFMA threads=%d: %.2f TFLOP/s (%.3f ms)\n", threads, fl / ms * 1e-9, ms);
    }
    for (int threads : {128, 256, 512, 1024}) {
        int iters = 20000;
        k_dmma884<8><<<sms, threads>>>(out, 100); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); k_dmma884<8><<<sms, threads>>>(out, iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 256 * 8 * iters * (double)(threads / 32) * sms;
        printf("DMMA m8n8k4 threads=%d: %.2f TFLOP/s (%.3f ms)\n", threads, fl / ms * 1e-9, ms);
    }
#ifdef BIGMMA
    for (int threads : {128, 256, 512, 1024}) {
        int iters = 20000;
        k_dmma1688<4><<<sms, threads>>>(out, 100); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); k_dmma1688<4><<<sms, threads>>>(out, iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 16 * 8 * 8 * 4 * iters * (double)(threads / 32) * sms;
        printf("DMMA m16n8k8 threads=%d: %.2f TFLOP/s (%.3f ms)\n", threads, fl / ms * 1e-9, ms);
    }
#endif
    for (int mode = 0; mode < 3; ++mode) {
        int iters = 20000, threads = 512;
        k_lds<<<sms, threads, 2048 * 16>>>(out, 100, mode); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); k_lds<<<sms, threads, 2048 * 16>>>(out, iters, mode); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        double n = 8.0 * iters * (threads / 32);   // warp-level LDS.128 per SM
        printf("LDS.128 mode=%d: %.2f warp-instr/us/SM  (%.3f ms)\n", mode, n / (ms * 1e3), ms);
    }
    return 0;
}
