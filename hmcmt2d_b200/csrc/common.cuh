// Common device helpers for the hmcmt_b200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace hmcmt {

constexpr double kMu0 = 4.0 * 3.14159265358979323846 * 1e-7;   // 4*pi*1e-7, mt1DField.jl:34
constexpr double kEps0 = 8.85 * 1e-12;                          // mt1DField.jl:35
constexpr double kPi = 3.14159265358979323846;

// Status codes follow the MUMPS convention the reference checks (MUMPSfuncs.jl:59-73).
enum Status : int {
    kOk = 0,
    kErrSingular = -10,
    kErrAlloc = -13,
    kErrNotPosDef = -40,
    kErrArg = -3,
    kErrBounds = -21,      // checkParameterBound!: a bound could not be met in 500 reflections (HMCSampler.jl:546-548)
    kErrCuda = -99,
    kErrNoDevice = -98,
};

#define HMCMT_CUDA_TRY(expr)                                                            \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            fprintf(stderr, "[hmcmt_b200] CUDA error %s at %s:%d: %s\n",                \
                    cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e));  \
            return (_e == cudaErrorMemoryAllocation) ? kErrAlloc : kErrCuda;            \
        }                                                                               \
    } while (0)

// ---- complex<double> as double2 -------------------------------------------------------
typedef double2 cplx;

__host__ __device__ __forceinline__ cplx mk(double r, double i) { return make_double2(r, i); }
__host__ __device__ __forceinline__ cplx operator+(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx operator-(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx operator-(cplx a) { return mk(-a.x, -a.y); }
__host__ __device__ __forceinline__ cplx operator*(cplx a, cplx b) {
    return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx operator*(double s, cplx a) { return mk(s * a.x, s * a.y); }
__host__ __device__ __forceinline__ cplx operator*(cplx a, double s) { return mk(s * a.x, s * a.y); }
__host__ __device__ __forceinline__ cplx operator/(cplx a, double s) { return mk(a.x / s, a.y / s); }
__host__ __device__ __forceinline__ cplx& operator+=(cplx& a, cplx b) { a.x += b.x; a.y += b.y; return a; }
__host__ __device__ __forceinline__ cplx& operator-=(cplx& a, cplx b) { a.x -= b.x; a.y -= b.y; return a; }
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return mk(a.x, -a.y); }
__host__ __device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }
__host__ __device__ __forceinline__ double cabs_(cplx a) { return hypot(a.x, a.y); }
// a += b*c
__host__ __device__ __forceinline__ void cfma(cplx& a, cplx b, cplx c) {
    a.x = fma(b.x, c.x, a.x); a.x = fma(-b.y, c.y, a.x);
    a.y = fma(b.x, c.y, a.y); a.y = fma(b.y, c.x, a.y);
}
// Smith's algorithm (what Julia / C99 use up to scaling): robust complex division.
__host__ __device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    if (fabs(b.x) >= fabs(b.y)) {
        double r = b.y / b.x, d = b.x + b.y * r;
        return mk((a.x + a.y * r) / d, (a.y - a.x * r) / d);
    } else {
        double r = b.x / b.y, d = b.x * r + b.y;
        return mk((a.x * r + a.y) / d, (a.y * r - a.x) / d);
    }
}
__host__ __device__ __forceinline__ cplx crecip(cplx b) { return cdiv(mk(1.0, 0.0), b); }
__host__ __device__ __forceinline__ cplx operator/(cplx a, cplx b) { return cdiv(a, b); }
__host__ __device__ __forceinline__ cplx operator/(double a, cplx b) { return cdiv(mk(a, 0.0), b); }

// principal square root
__host__ __device__ inline cplx csqrt_(cplx z) {
    double a = z.x, b = z.y;
    if (a == 0.0 && b == 0.0) return mk(0.0, b);
    double m = hypot(a, b);
    double t = sqrt(0.5 * (m + fabs(a)));
    if (a >= 0.0) return mk(t, b / (2.0 * t));
    return mk(fabs(b) / (2.0 * t), copysign(t, b));
}
__host__ __device__ inline cplx cexp_(cplx z) {
    double e = exp(z.x), s, c;
    sincos(z.y, &s, &c);
    return mk(e * c, e * s);
}
// tanh(z) with overflow-safe branch for large |Re z| (matches libm / Julia to round-off)
__host__ __device__ inline cplx ctanh_(cplx z) {
    double x = z.x, y = z.y;
    if (fabs(x) > 22.0) {
        // tanh(x+iy) -> sign(x) + i * 4 sin(y) cos(y) exp(-2|x|)
        double s, c;
        sincos(y, &s, &c);
        return mk(copysign(1.0, x), 4.0 * s * c * exp(-2.0 * fabs(x)));
    }
    double t = tan(y), beta = 1.0 + t * t;        // sec^2 y
    double sh = sinh(x), rho = sqrt(1.0 + sh * sh);   // cosh x
    if (isinf(t)) return mk(rho / sh, 1.0 / t);
    double den = 1.0 + beta * sh * sh;
    return mk(beta * rho * sh / den, t / den);
}

// ---- FP64 tensor-core MMA: D(8x8) += A(8x4) * B(4x8)  (SASS: DMMA.8x8x4) ---------------
// Fragment layout (PTX ISA, m8n8k4 .f64): lane = 4*g + t;  A[g][t], B[t][g], C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}

// same, predicated on a warp-uniform flag: no branch, so independent MMAs can be scheduled across tiles
__device__ __forceinline__ void dmma884_p(double (&c)[2], double a, double b, int pred) {
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %4, 0;\n\t"
        "@q mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n\t}"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b), "r"(pred));
}

// ---- named barriers -------------------------------------------------------------------
// bar.sync / bar.arrive must be executed by a converged warp: reconverge explicitly first
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    __syncwarp();
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
    __syncwarp();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// CTA barrier preceded by an explicit warp reconvergence: bar.sync is an *aligned* barrier, a warp that reaches it
// in two divergent halves is counted twice and deadlocks the next generation (seen with ncu/trace on B200).
__device__ __forceinline__ void cta_sync() {
    __syncwarp();
    __syncthreads();
}

// ---- TMA bulk copies (SASS: UBLKCP) + mbarrier ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() {
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy, completion on mbarrier (bytes multiple of 16, 16B aligned)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 16-byte asynchronous global -> shared copy (LDGSTS); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// global -> L2 prefetch of a contiguous range (bytes multiple of 16, 16B aligned); no completion tracking
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace hmcmt
