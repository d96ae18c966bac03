// Forward + backward substitution with the block-LDL^T factor written by band_factor_kernel.
// Replaces `applyMUMPS(Ainv, rhs)` / `Ainv \ rhs` for the adjoint solve (compJacTMatVec.jl:220-224,
// 291-295) and `solve_mumps_cmplx_` (MUMPSfuncs.jl:123-132).  HBM-bound: streams the 16*8T*8-byte
// panel images with TMA bulk loads (kSolveStages panels in flight on mbarriers) — 2 reads of the
// factor per right-hand side.
#pragma once
#include "band_factor.cuh"

namespace hmcmt {

struct SolveJob {
    const double* panels;   // factor of the system
    const cplx* ainvz;      // [S][72] (A11^{-1} | z of the fused system, unused here)
    const cplx* rhs;        // [N] internal ordering
    cplx* x;                // [N] (may alias rhs)
    cplx* zbuf;             // [S*8] scratch
};

constexpr int kSolveStages = 8;
constexpr int kSolveThreads = 256;

template <int T>
struct SolveSmem {
    static constexpr int R = TS * T;
    double stage[kSolveStages][2][2][R][4];
    cplx ainv[kSolveStages][64];
    cplx y[R];
    cplx zv[8];
    cplx part[kSolveThreads / 32][8];
    cplx ringRhs[kRing][8];     // rhs rows entering the window (forward sweep), prefetched kPre steps ahead
    cplx ringZ[kRing][8];       // z of upcoming panels (backward sweep)
    uint64_t mbar[kSolveStages];
};

template <int T>
__global__ void __launch_bounds__(kSolveThreads, 1)
band_solve_kernel(const SolveJob* __restrict__ jobs, int N) {
    constexpr int R = TS * T, NTHR = kSolveThreads, NST = kSolveStages;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SolveSmem<T>& sm = *reinterpret_cast<SolveSmem<T>*>(smem_raw);
    const SolveJob job = jobs[blockIdx.x];
    const int S = (N + TS - 1) / TS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t PBYTES = panel_doubles(T) * 8;

    for (int i = tid; i < R; i += NTHR) sm.y[i] = (i < N) ? job.rhs[i] : mk(0.0, 0.0);
    if (tid == 0) {
        for (int q = 0; q < NST; ++q) mbar_init(&sm.mbar[q], 1);
        fence_mbar_init();
    }
    __syncthreads();
    // global iteration counter `it` runs over the 2S panel visits (forward then backward)
    auto panel_of = [&](int it) { return it < S ? it : 2 * S - 1 - it; };
    auto issue = [&](int it) {
        int s = panel_of(it), st = it % NST;
        mbar_arrive_expect_tx(&sm.mbar[st], PBYTES + 64 * 16);
        bulk_g2s(&sm.stage[st][0][0][0][0], job.panels + (size_t)s * panel_doubles(T), PBYTES, &sm.mbar[st]);
        bulk_g2s(&sm.ainv[st][0], job.ainvz + (size_t)s * AZ, 64 * 16, &sm.mbar[st]);
    };
    if (tid == 0)
        for (int q = 0; q < NST - 1 && q < 2 * S; ++q) issue(q);
    if (tid >= 32 && tid < 40)
        for (int q = 0; q < kPre; ++q) {
            int gnew = (q + T) * TS + (tid - 32);
            sm.ringRhs[q % kRing][tid - 32] = (gnew < N) ? job.rhs[gnew] : mk(0.0, 0.0);
        }
    __syncthreads();

    // ---------------- forward:  z_s = A11^{-1} y_p ;  y_rest -= raw_s z_s ----------------
    for (int it = 0; it < S; ++it) {
        const int s = it, p = s % T, st = it % NST, rp = p * TS;
        if (tid == 0 && it + NST - 1 < 2 * S) issue(it + NST - 1);
        cplx pre = mk(0.0, 0.0);
        if (tid >= 32 && tid < 40) {          // rhs rows of the block entering kPre steps from now (load in flight over the step)
            int gnew = (s + kPre + T) * TS + (tid - 32);
            if (gnew < N) pre = job.rhs[gnew];
        }
        mbar_wait(&sm.mbar[st], (uint32_t)((it / NST) & 1));
        if (warp == 0) {
            const int ii = lane >> 2, tt = lane & 3;
            cplx acc = sm.ainv[st][ii * 8 + 2 * tt] * sm.y[rp + 2 * tt] + sm.ainv[st][ii * 8 + 2 * tt + 1] * sm.y[rp + 2 * tt + 1];
#pragma unroll
            for (int off = 1; off <= 2; off <<= 1) {
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
                acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
            }
            if (tt == 0) { sm.zv[ii] = acc; job.zbuf[(size_t)s * 8 + ii] = acc; }
        }
        __syncthreads();
        for (int r = tid; r < R; r += NTHR) {
            if ((r >> 3) == p) {      // recycle: slot block p now holds global block s+T
                sm.y[r] = sm.ringRhs[s % kRing][r & 7];
                continue;
            }
            cplx acc = sm.y[r];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                cplx rv = mk(sm.stage[st][0][k >> 2][r][k & 3], sm.stage[st][1][k >> 2][r][k & 3]);
                cfma(acc, -rv, sm.zv[k]);
            }
            sm.y[r] = acc;
        }
        if (tid >= 32 && tid < 40) sm.ringRhs[(s + kPre) % kRing][tid - 32] = pre;
        __syncthreads();
    }
    // ---------------- backward:  x_p = z_s - A11^{-1} raw_s^T x_rest ----------------
    for (int i = tid; i < R; i += NTHR) sm.y[i] = mk(0.0, 0.0);
    __threadfence();              // zbuf was written by this CTA during the forward sweep
    __syncthreads();
    if (tid >= 32 && tid < 40)
        for (int q = 0; q < kPre && S - 1 - q >= 0; ++q) sm.ringZ[q % kRing][tid - 32] = job.zbuf[(size_t)(S - 1 - q) * 8 + (tid - 32)];
    __syncthreads();
    constexpr int NRG = NTHR / 8;
    for (int it = S; it < 2 * S; ++it) {
        const int s = panel_of(it), p = s % T, st = it % NST;
        if (tid == 0 && it + NST - 1 < 2 * S) issue(it + NST - 1);
        const int jb = it - S;                 // backward step counter
        cplx prez = mk(0.0, 0.0);
        if (tid >= 32 && tid < 40 && s - kPre >= 0) prez = job.zbuf[(size_t)(s - kPre) * 8 + (tid - 32)];
        mbar_wait(&sm.mbar[st], (uint32_t)((it / NST) & 1));
        const int c = tid & 7, rg = tid >> 3;
        cplx acc = mk(0.0, 0.0);
        for (int r = rg; r < R; r += NRG) {
            if ((r >> 3) == p) continue;
            cplx rv = mk(sm.stage[st][0][c >> 2][r][c & 3], sm.stage[st][1][c >> 2][r][c & 3]);
            cfma(acc, rv, sm.y[r]);
        }
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
        if (lane < 8) sm.part[warp][lane] = acc;
        __syncthreads();
        if (warp == 0) {
            cplx d = mk(0.0, 0.0);
            for (int w = (lane >> 3); w < NTHR / 32; w += 4) d += sm.part[w][lane & 7];
#pragma unroll
            for (int off = 8; off <= 16; off <<= 1) {
                d.x += __shfl_xor_sync(0xffffffffu, d.x, off);
                d.y += __shfl_xor_sync(0xffffffffu, d.y, off);
            }
            const int ii = lane >> 2, tt = lane & 3;
            cplx d0 = mk(__shfl_sync(0xffffffffu, d.x, 2 * tt), __shfl_sync(0xffffffffu, d.y, 2 * tt));
            cplx d1 = mk(__shfl_sync(0xffffffffu, d.x, 2 * tt + 1), __shfl_sync(0xffffffffu, d.y, 2 * tt + 1));
            cplx xv = sm.ainv[st][ii * 8 + 2 * tt] * d0 + sm.ainv[st][ii * 8 + 2 * tt + 1] * d1;
#pragma unroll
            for (int off = 1; off <= 2; off <<= 1) {
                xv.x += __shfl_xor_sync(0xffffffffu, xv.x, off);
                xv.y += __shfl_xor_sync(0xffffffffu, xv.y, off);
            }
            if (tt == 0) {
                cplx xo = sm.ringZ[jb % kRing][ii] - xv;
                sm.y[p * TS + ii] = xo;
                int gidx = s * TS + ii;
                if (gidx < N) job.x[gidx] = xo;
            }
        }
        if (tid >= 32 && tid < 40) sm.ringZ[(jb + kPre) % kRing][tid - 32] = prez;
        __syncthreads();
    }
}

}  // namespace hmcmt
