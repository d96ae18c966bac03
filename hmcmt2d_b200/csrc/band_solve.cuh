// Forward + backward substitution with the block-LDL^T factor written by band_factor_kernel.
// Replaces `applyMUMPS(Ainv, rhs)` / `Ainv \ rhs` for the adjoint solve (compJacTMatVec.jl:220-224,
// 291-295) and `solve_mumps_cmplx_` (MUMPSfuncs.jl:123-132).  HBM-bound: streams the 16*8T*8-byte
// panel images with TMA bulk loads (double-buffered on mbarriers) — 2 reads of the factor per rhs.
#pragma once
#include "band_factor.cuh"

namespace hmcmt {

struct SolveJob {
    const double* panels;   // factor of the system
    const cplx* ainv;
    const cplx* rhs;        // [N] internal ordering
    cplx* x;                // [N] (may alias rhs)
    cplx* zbuf;             // [S*8] scratch
};

template <int T>
struct SolveSmem {
    static constexpr int R = TS * T;
    double raw[2][2][2][R][4];
    cplx ainv[2][64];
    cplx y[R];
    cplx zv[8];
    cplx part[8][8];
    cplx dots[8];
    uint64_t mbar[2];
};

constexpr int kSolveThreads = 256;

template <int T>
__global__ void __launch_bounds__(kSolveThreads, 1)
band_solve_kernel(const SolveJob* __restrict__ jobs, int N) {
    constexpr int R = TS * T, NTHR = kSolveThreads;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SolveSmem<T>& sm = *reinterpret_cast<SolveSmem<T>*>(smem_raw);
    const SolveJob job = jobs[blockIdx.x];
    const int S = (N + TS - 1) / TS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t PBYTES = panel_doubles(T) * 8;

    for (int i = tid; i < R; i += NTHR) sm.y[i] = (i < N) ? job.rhs[i] : mk(0.0, 0.0);
    if (tid == 0) {
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int s) {
        int bf = s & 1;
        mbar_arrive_expect_tx(&sm.mbar[bf], PBYTES + 64 * 16);
        bulk_g2s(&sm.raw[bf][0][0][0][0], job.panels + (size_t)s * panel_doubles(T), PBYTES, &sm.mbar[bf]);
        bulk_g2s(&sm.ainv[bf][0], job.ainv + (size_t)s * 64, 64 * 16, &sm.mbar[bf]);
    };
    uint32_t phase[2] = {0, 0};

    // ---------------- forward:  z_s = A11^{-1} y_p ;  y_rest -= raw_s z_s ----------------
    if (tid == 0) issue(0);
    for (int s = 0; s < S; ++s) {
        const int p = s % T, bf = s & 1, rp = p * TS;
        if (tid == 0 && s + 1 < S) issue(s + 1);
        mbar_wait(&sm.mbar[bf], phase[bf]);
        phase[bf] ^= 1;
        if (tid < 8) {
            cplx acc = mk(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < 8; ++k) cfma(acc, sm.ainv[bf][tid * 8 + k], sm.y[rp + k]);
            sm.zv[tid] = acc;
            job.zbuf[(size_t)s * 8 + tid] = acc;
        }
        __syncthreads();
        for (int r = tid; r < R; r += NTHR) {
            if ((r >> 3) == p) {      // recycle: slot block p now holds global block s+T
                int gnew = (s + T) * TS + (r & 7);
                sm.y[r] = (gnew < N) ? job.rhs[gnew] : mk(0.0, 0.0);
                continue;
            }
            cplx acc = sm.y[r];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                cplx rv = mk(sm.raw[bf][0][k >> 2][r][k & 3], sm.raw[bf][1][k >> 2][r][k & 3]);
                cfma(acc, -rv, sm.zv[k]);
            }
            sm.y[r] = acc;
        }
        __syncthreads();
    }
    // ---------------- backward:  x_p = z_s - A11^{-1} raw_s^T x_rest ----------------
    for (int i = tid; i < R; i += NTHR) sm.y[i] = mk(0.0, 0.0);
    __syncthreads();
    if (tid == 0) issue(S - 1);          // buffer (S-1)&1 was last consumed at s=S-1 (all threads past the sync)
    // NB parity bookkeeping continues: buffer (S-1)&1 is reused immediately, its phase already toggled.
    constexpr int NRG = NTHR / 8;
    for (int s = S - 1; s >= 0; --s) {
        const int p = s % T, bf = s & 1;
        if (tid == 0 && s > 0) issue(s - 1);
        mbar_wait(&sm.mbar[bf], phase[bf]);
        phase[bf] ^= 1;
        const int c = tid & 7, rg = tid >> 3;
        cplx acc = mk(0.0, 0.0);
        for (int r = rg; r < R; r += NRG) {
            if ((r >> 3) == p) continue;
            cplx rv = mk(sm.raw[bf][0][c >> 2][r][c & 3], sm.raw[bf][1][c >> 2][r][c & 3]);
            cfma(acc, rv, sm.y[r]);
        }
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
        if (lane < 8) sm.part[warp][lane] = acc;
        __syncthreads();
        if (tid < 8) {
            cplx d = mk(0.0, 0.0);
#pragma unroll
            for (int w = 0; w < NTHR / 32; ++w) d += sm.part[w][tid];
            sm.dots[tid] = d;
        }
        __syncthreads();
        if (tid < 8) {
            cplx xv = job.zbuf[(size_t)s * 8 + tid];
#pragma unroll
            for (int k = 0; k < 8; ++k) cfma(xv, -sm.ainv[bf][tid * 8 + k], sm.dots[k]);
            sm.y[p * TS + tid] = xv;
            int gidx = s * TS + tid;
            if (gidx < N) job.x[gidx] = xv;
        }
        __syncthreads();
    }
}

}  // namespace hmcmt
