// Forward + backward substitution with the block-LDL^T factor written by band_factor_kernel.
// Replaces `applyMUMPS(Ainv, rhs)` / `Ainv \ rhs` for the adjoint solve (compJacTMatVec.jl:220-224,
// 291-295) and `solve_mumps_cmplx_` (MUMPSfuncs.jl:123-132).  HBM-bound: streams the 16*8T*8-byte
// panel images with TMA bulk loads (solve_stages(T) panels in flight on mbarriers) — 2 reads of the
// factor per right-hand side.  Split systems use the same launch sequence as the factorisation
// (FM_OWN: forward sweep of both halves, FM_SEP: separator forward + backward, FM_BACK: backward
// sweep of both halves), hand-over through global scratch in stream order.
#pragma once
#include "band_factor.cuh"

namespace hmcmt {

struct SolveJob {
    const double* panels[2];   // factor of the system, per half (rank)
    const cplx* ainvz[2];      // [steps][72] (A11^{-1} | z of the fused system, unused here)
    const cplx* rhs;           // [N] internal ordering
    cplx* x;                   // [N] (may alias rhs)
    cplx* zbuf[2];             // [steps*8] scratch per rank
    cplx* wexp;                // split only: hand-over scratch (the [2][R] rhs windows and [R] separator solution after the window images)
};

// TMA stages in flight: 8 for the register-window sizes (14 KB panels), fewer for the large-bandwidth windows (up to 45 KB)
__host__ __device__ constexpr int solve_stages(int T) { return T <= 14 ? 8 : (T <= 28 ? 5 : 4); }
constexpr int kSolveThreads = 256;
// extra launch mode of the solve kernel: backward sweep only, z read from the factor's [A11^{-1} | z] stream (the fused
// forward system of the large-bandwidth factorisation, band_big.cuh)
constexpr int SM_BACKZ = 4;

template <int T>
struct SolveSmem {
    static constexpr int R = TS * T;
    static constexpr int NST = solve_stages(T);
    double stage[NST][2][2][R][4];
    cplx ainv[NST][64];
    cplx y[R];
    cplx zv[8];
    cplx part[kSolveThreads / 32][8];
    cplx ringRhs[kRing][8];     // rhs rows entering the window (forward sweep), prefetched kPre steps ahead
    cplx ringZ[kRing][8];       // z of upcoming panels (backward sweep)
    uint64_t mbar[NST];
};

__device__ __forceinline__ cplx local_rhs(const LocalDom& L, const cplx* rhs, int g) {
    int kind, lrel;
    const int q = L.map(g, kind, lrel);
    if (q < 0 || (kind == 1 && L.rank == 1)) return mk(0.0, 0.0);
    return rhs[q];
}
__device__ __forceinline__ void local_store(const LocalDom& L, cplx* x, int g, cplx v) {
    int kind, lrel;
    const int q = L.map(g, kind, lrel);
    if (q >= 0 && !(kind == 1 && L.rank == 1)) x[q] = v;
}

template <int T>
__global__ void __launch_bounds__(kSolveThreads, 1)
band_solve_kernel(const SolveJob* __restrict__ jobs, BandDom dom, int mode) {
    constexpr int R = TS * T, NTHR = kSolveThreads, NST = solve_stages(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SolveSmem<T>& sm = *reinterpret_cast<SolveSmem<T>*>(smem_raw);
    const bool paired = (mode == FM_OWN || mode == FM_BACK);
    const int rank = paired ? (int)(blockIdx.x & 1) : 0;
    const SolveJob job = jobs[paired ? (blockIdx.x >> 1) : blockIdx.x];
    const LocalDom L = LocalDom::make(dom, rank);
    const double* const panels = job.panels[rank];
    const cplx* const ainvz = job.ainvz[rank];
    cplx* const zbuf = job.zbuf[rank];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int nLoc = L.nLoc;
    constexpr uint32_t PBYTES = panel_doubles(T) * 8;
    const int sBeg = (mode == FM_SEP) ? L.sOwn : 0;
    cplx* const yimg0 = job.wexp + 2 * (size_t)R * R;
    cplx* const yimg1 = yimg0 + R;
    cplx* const xsep = yimg0 + 2 * R;
    auto rel = [&](int slot, int base) { int a = slot - base % T; return a < 0 ? a + T : a; };

    // rhs may alias x: all rhs reads happen in the FM_OWN / FM_FULL forward sweep (and, for the separator rows, come through
    // the exported windows), all x writes of a split system in the later FM_SEP / FM_BACK launches.
    if (mode != FM_BACK && mode != SM_BACKZ)
        for (int i_b = 0; i_b < R; i_b += NTHR) if (const int i = i_b + tid; i < R) {
            cplx v = mk(0.0, 0.0);
            if (mode == FM_SEP) {
                const int a = rel(i >> 3, sBeg) * TS + (i & 7);
                v = yimg0[a] + yimg1[a];
            } else if (i < nLoc) v = local_rhs(L, job.rhs, i);
            sm.y[i] = v;
        }
    if (tid == 0) {
        for (int q = 0; q < NST; ++q) mbar_init(&sm.mbar[q], 1);
        fence_mbar_init();
    }
    if (tid >= 32 && tid < 40 && mode != FM_BACK && mode != SM_BACKZ)
        for (int q = 0; q < kPre; ++q) {
            int gnew = (sBeg + q + T) * TS + (tid - 32);
            sm.ringRhs[(sBeg + q) % kRing][tid - 32] = (gnew < nLoc) ? local_rhs(L, job.rhs, gnew) : mk(0.0, 0.0);
        }
    cta_sync();
    int itBase = 0;        // running count of panel visits: stage = visit % NST, parity = (visit / NST) & 1
    auto issue = [&](int s, int visit) {
        const int st = visit % NST;
        mbar_arrive_expect_tx(&sm.mbar[st], PBYTES + 64 * 16);
        bulk_g2s(&sm.stage[st][0][0][0][0], panels + (size_t)s * panel_doubles(T), PBYTES, &sm.mbar[st]);
        bulk_g2s(&sm.ainv[st][0], ainvz + (size_t)s * AZ, 64 * 16, &sm.mbar[st]);
    };

    // ---------------- forward over steps [sLo, sHi):  z_s = A11^{-1} y_p ;  y_rest -= raw_s z_s ----------------
    auto forward_range = [&](int sLo, int sHi) {
        const int n = sHi - sLo;
        if (tid == 0)
            for (int k = 0; k < NST - 1 && k < n; ++k) issue(sLo + k, itBase + k);
        for (int k = 0; k < n; ++k) {
            const int s = sLo + k, p = s % T, it = itBase + k, st = it % NST, rp = p * TS;
            if (tid == 0 && k + NST - 1 < n) issue(s + NST - 1, it + NST - 1);
            cplx pre = mk(0.0, 0.0);
            if (tid >= 32 && tid < 40) {          // rhs rows of the block entering kPre steps from now (load in flight over the step)
                int gnew = (s + kPre + T) * TS + (tid - 32);
                if (gnew < nLoc) pre = local_rhs(L, job.rhs, gnew);
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
                if (tt == 0) { sm.zv[ii] = acc; zbuf[(size_t)s * 8 + ii] = acc; }
            }
            cta_sync();
            for (int r_b = 0; r_b < R; r_b += NTHR) if (const int r = r_b + tid; r < R) {
                if ((r >> 3) == p) {      // recycle: slot block p now holds local block s+T
                    sm.y[r] = sm.ringRhs[s % kRing][r & 7];
                    continue;
                }
                cplx acc = sm.y[r];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    cplx rv = mk(sm.stage[st][0][kk >> 2][r][kk & 3], sm.stage[st][1][kk >> 2][r][kk & 3]);
                    cfma(acc, -rv, sm.zv[kk]);
                }
                sm.y[r] = acc;
            }
            if (tid >= 32 && tid < 40) sm.ringRhs[(s + kPre) % kRing][tid - 32] = pre;
            cta_sync();
        }
        itBase += n;
    };
    // ---------------- backward over steps sHi-1 .. sLo:  x_p = z_s - A11^{-1} raw_s^T x_rest ----------------
    constexpr int NRG = NTHR / 8;
    auto zsrc = [&](int s, int i) { return mode == SM_BACKZ ? ainvz[(size_t)s * AZ + 64 + i] : zbuf[(size_t)s * 8 + i]; };
    auto backward_range = [&](int sHi, int sLo) {
        const int n = sHi - sLo;
        if (tid == 0)
            for (int k = 0; k < NST - 1 && k < n; ++k) issue(sHi - 1 - k, itBase + k);
        if (tid >= 32 && tid < 40)
            for (int q = 0; q < kPre && q < n; ++q) sm.ringZ[q % kRing][tid - 32] = zsrc(sHi - 1 - q, tid - 32);
        cta_sync();
        for (int k = 0; k < n; ++k) {
            const int s = sHi - 1 - k, p = s % T, it = itBase + k, st = it % NST;
            if (tid == 0 && k + NST - 1 < n) issue(s - (NST - 1), it + NST - 1);
            cplx prez = mk(0.0, 0.0);
            if (tid >= 32 && tid < 40 && k + kPre < n) prez = zsrc(s - kPre, tid - 32);
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
            cta_sync();
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
                    cplx xo = sm.ringZ[k % kRing][ii] - xv;
                    sm.y[p * TS + ii] = xo;
                    local_store(L, job.x, s * TS + ii, xo);
                }
            }
            if (tid >= 32 && tid < 40) sm.ringZ[(k + kPre) % kRing][tid - 32] = prez;
            cta_sync();
        }
        itBase += n;
    };

    if (mode == FM_OWN) {
        forward_range(0, L.sOwn);
        cplx* const yimg = rank == 0 ? yimg0 : yimg1;
        for (int r_b = 0; r_b < R; r_b += NTHR) if (const int r = r_b + tid; r < R) yimg[rel(r >> 3, L.sOwn) * TS + (r & 7)] = sm.y[r];
        return;
    }
    if (mode == FM_FULL) forward_range(0, L.sTot);
    if (mode == FM_SEP) forward_range(L.sOwn, L.sTot);
    __threadfence();              // zbuf was written by this CTA during the forward sweep (earlier launches are ordered by the stream)
    cta_sync();
    for (int i_b = 0; i_b < R; i_b += NTHR) if (const int i = i_b + tid; i < R)
        sm.y[i] = (mode == FM_BACK) ? xsep[rel(i >> 3, L.sOwn) * TS + (i & 7)] : mk(0.0, 0.0);
    cta_sync();
    if (mode == FM_FULL || mode == SM_BACKZ) {
        backward_range(L.sTot, 0);
    } else if (mode == FM_SEP) {
        backward_range(L.sTot, L.sOwn);
        for (int r_b = 0; r_b < R; r_b += NTHR) if (const int r = r_b + tid; r < R) xsep[rel(r >> 3, L.sOwn) * TS + (r & 7)] = sm.y[r];
    } else {
        backward_range(L.sOwn, 0);
    }
}

}  // namespace hmcmt
