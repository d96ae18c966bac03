// Forward + backward substitution with the block-LDL^T factor written by band_factor_kernel.
// Replaces `applyMUMPS(Ainv, rhs)` / `Ainv \ rhs` for the adjoint solve (compJacTMatVec.jl:220-224,
// 291-295), the back-substitution half of the forward solve (mt2DTE.jl:53, mt2DTM.jl:52) and `solve_mumps_cmplx_`
// (MUMPSfuncs.jl:123-132).  HBM-bound: streams the 16*8T*8-byte panel images with TMA bulk loads (solve_stages(T) panels in
// flight on mbarriers), one read of the factor per sweep.  Each sweep is a software pipeline of three warp roles (near /
// far / service, see below) so that the dependency between consecutive 8-column steps runs inside one warp.
// Split systems use the same launch sequence as the factorisation (FM_OWN: forward sweep of both halves, FM_SEP: separator
// forward + backward, FM_BACK: backward sweep of both halves), hand-over through global scratch in stream order.
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
constexpr int kSolveRing = 8, kRhsAhead = 5;   // rhs ring of the forward sweep: depth and prefetch distance (steps)
enum { SB_X0 = 1, SB_X1 = 2, SB_F0 = 3, SB_F1 = 4 };      // named barriers of the sweep pipeline
// extra launch mode of the solve kernel: backward sweep only, z read from the factor's [A11^{-1} | z] stream (the fused
// forward system of a split factorisation)
constexpr int SM_BACKZ = 4;
// the same for the two halves of a split system (2 CTAs per system, window initialised with the separator solution)
constexpr int SM_BACKZ_OWN = 5;

template <int T>
struct SolveSmem {
    static constexpr int R = TS * T;
    static constexpr int NST = solve_stages(T);
    double stage[NST][2][2][R][4];
    cplx az[NST][AZ];                           // per stage: A11^{-1} (64) and, for the backward sweep, z (8)
    cplx y[R];
    cplx zv[2][8];                              // z of the current step, double-buffered by step parity
    cplx part[2][kSolveThreads / 32][8];        // far warps' partial dot products, double-buffered by step parity
    cplx ringRhs[kSolveRing][8];                // rhs rows entering the window (forward sweep), cp.async'ed kRhsAhead steps ahead
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
    const bool paired = (mode == FM_OWN || mode == FM_BACK || mode == SM_BACKZ_OWN);
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
    if (mode != FM_BACK && mode != SM_BACKZ && mode != SM_BACKZ_OWN)
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
    cta_sync();
    int itBase = 0;        // running count of panel visits: stage = visit % NST, parity = (visit / NST) & 1
    auto issue = [&](int s, int visit, int dir) {      // dir: +1 forward sweep, -1 backward sweep
        const int st = visit % NST;
        // backward sweep: z_s travels with its panel (from the factor's [A11^{-1} | z] stream, or from this solve's zbuf)
        const bool zFromFactor = (mode == SM_BACKZ || mode == SM_BACKZ_OWN);
        const uint32_t azBytes = (dir < 0 && zFromFactor) ? AZ * 16 : 64 * 16;
        mbar_arrive_expect_tx(&sm.mbar[st], PBYTES + azBytes + ((dir < 0 && !zFromFactor) ? 8 * 16 : 0));
        bulk_g2s(&sm.stage[st][0][0][0][0], panels + (size_t)s * panel_doubles(T), PBYTES, &sm.mbar[st]);
        bulk_g2s(&sm.az[st][0], ainvz + (size_t)s * AZ, azBytes, &sm.mbar[st]);
        if (dir < 0 && !zFromFactor) bulk_g2s(&sm.az[st][64], zbuf + (size_t)s * 8, 8 * 16, &sm.mbar[st]);
    };

    // Both sweeps are software-pipelined across two warp roles so that the serial dependency between consecutive 8-column
    // steps runs inside ONE warp (no CTA-wide barrier on the chain):
    //   near warp (warp 0)  : the 8x8 work that links step s to step s+1 (pivot block and the neighbouring block);
    //   far warps (1..6)    : the remaining rows of the panel, one step out of phase with the near warp;
    //   service warp (7)    : global stores, ring prefetches and TMA issue.
    // Named barriers, two per direction alternating with the step parity: SB_X* near -> far, SB_F* far -> near.
    constexpr int NFAR = NTHR - 32;
    auto reduce_rows_in_warp = [&](cplx& acc) {      // lanes differing in bits 3,4 hold partial sums of the same column
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
    };
    auto reduce_quad = [&](cplx& acc) {               // the 4 lanes of a row
#pragma unroll
        for (int off = 1; off <= 2; off <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
    };

    // ---------------- forward over steps [sLo, sHi):  z_s = A11^{-1} y_p ;  y_rest -= raw_s z_s ----------------
    auto forward_range = [&](int sLo, int sHi) {
        const int n = sHi - sLo;
        if (n <= 0) return;
        if (tid == 0)
            for (int k = 0; k < NST && k < n; ++k) issue(sLo + k, itBase + k, +1);
        if (warp == 0) {
            // the near warp touches shared memory only: global stores, ring prefetches and TMA issue belong to the far warps
            const int ii = lane >> 2, tt = lane & 3;
            for (int k = 0; k < n; ++k) {
                const int s = sLo + k, p = s % T, p1 = (s + 1) % T, it = itBase + k, st = it % NST, rp = p * TS;
                mbar_wait(&sm.mbar[st], (uint32_t)((it / NST) & 1));
                // z_s = A11^{-1} y_p  (y_p is final: far(s-2) was joined in the previous iteration, the block update of s-1 is ours)
                cplx z = sm.az[st][ii * 8 + 2 * tt] * sm.y[rp + 2 * tt] + sm.az[st][ii * 8 + 2 * tt + 1] * sm.y[rp + 2 * tt + 1];
                reduce_quad(z);
                if (tt == 0) sm.zv[k & 1][ii] = z;
                bar_arrive(SB_X0 + (k & 1), NTHR);                       // z_s published: far(s) may start (bar.arrive orders the smem writes)
                // block s+1:  y_{p1} -= raw_s[p1 rows] z_s   — after far(s-1) has finished with those rows
                const int r1 = p1 * TS + ii;
                cplx u = mk(0.0, 0.0);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = 2 * tt + e;
                    const cplx rv = mk(sm.stage[st][0][c >> 2][r1][c & 3], sm.stage[st][1][c >> 2][r1][c & 3]);
                    const cplx zc = mk(__shfl_sync(0xffffffffu, z.x, 4 * c), __shfl_sync(0xffffffffu, z.y, 4 * c));
                    cfma(u, rv, zc);
                }
                reduce_quad(u);
                if (k >= 1) bar_sync(SB_F0 + ((k - 1) & 1), NTHR);
                if (tt == 0 && T > 1) sm.y[r1] = sm.y[r1] - u;
                __syncwarp();
            }
            bar_sync(SB_F0 + ((n - 1) & 1), NTHR);
        } else {
            // far warps (1..6) and the service warp (7) share one loop, so that every named barrier has a single waiting and
            // a single arriving instruction.
            // service warp: everything that touches global memory, kept off both the near chain and the far arithmetic.
            // lanes 0..7 store z_s, lane 8 refills the TMA stage of step s-1, lanes 16..23 feed the rhs ring with cp.async
            // issued kRhsAhead steps before the rows are needed, so the warp never waits on a load
            const bool service = warp == NTHR / 32 - 1;
            const int ft = tid - 32, rl = lane - 16;
            constexpr int NFC = NTHR - 64;         // far compute threads: one per window row
            auto rhs_fetch = [&](int step) {     // rhs rows of the block entering the window at `step` -> ring, asynchronously
                if (rl >= 0 && rl < 8) {
                    int kind, lrel;
                    const int gnew = (step + T) * TS + rl;
                    const int qg = (step < sHi && gnew < nLoc) ? L.map(gnew, kind, lrel) : -1;
                    const bool ok = qg >= 0 && !(kind == 1 && L.rank == 1);
                    cp_async16(&sm.ringRhs[step % kSolveRing][rl], job.rhs + (ok ? qg : 0), ok ? 16u : 0u);
                }
                cp_async_commit();
            };
            if (service) {
                for (int q = 0; q < kRhsAhead; ++q) rhs_fetch(sLo + q);
                cp_async_wait<kRhsAhead - 1>();
            }
            for (int k = 0; k < n; ++k) {
                const int s = sLo + k, p = s % T, p1 = (s + 1) % T, it = itBase + k, st = it % NST;
                if (!service) mbar_wait(&sm.mbar[st], (uint32_t)((it / NST) & 1));
                bar_sync(SB_X0 + (k & 1), NTHR);                         // z_s ready
                if (service) {
                    if (lane < 8) zbuf[(size_t)s * 8 + lane] = sm.zv[k & 1][lane];
                    // the stage of step s-1 is free: near is past it and far(s-1) arrived on SB_F in the previous iteration
                    if (lane == 8 && k >= 1 && k - 1 + NST < n) issue(s - 1 + NST, it - 1 + NST, +1);
                    rhs_fetch(s + kRhsAhead);
                    cp_async_wait<kRhsAhead - 1>();                       // the rows of step s+1 have landed
                } else {
#pragma unroll
                    for (int r_b = 0; r_b < R; r_b += NFC) {
                        const int r = r_b + ft;
                        const int slot = r >> 3;
                        if (r < R && slot == p) sm.y[r] = sm.ringRhs[s % kSolveRing][r & 7];      // recycle: slot p now holds local block s+T
                        if (r < R && slot != p && slot != p1) {                                   // (p1 is the near warp's block)
                            cplx a0 = sm.y[r], a1 = mk(0.0, 0.0);
#pragma unroll
                            for (int c = 0; c < 8; c += 2) {
                                cfma(a0, -mk(sm.stage[st][0][c >> 2][r][c & 3], sm.stage[st][1][c >> 2][r][c & 3]), sm.zv[k & 1][c]);
                                cfma(a1, -mk(sm.stage[st][0][c >> 2][r][(c & 3) + 1], sm.stage[st][1][c >> 2][r][(c & 3) + 1]), sm.zv[k & 1][c + 1]);
                            }
                            sm.y[r] = a0 + a1;
                        }
                    }
                }
                bar_arrive(SB_F0 + (k & 1), NTHR);
            }
        }
        cta_sync();
        itBase += n;
    };
    // ---------------- backward over steps sHi-1 .. sLo:  x_p = z_s - A11^{-1} raw_s^T x_rest ----------------
    auto backward_range = [&](int sHi, int sLo) {
        const int n = sHi - sLo;
        if (n <= 0) return;
        if (tid == 0)
            for (int k = 0; k < NST && k < n; ++k) issue(sHi - 1 - k, itBase + k, -1);
        cta_sync();
        if (warp == 0) {
            const int c = lane & 7, q = lane >> 3;
            for (int k = 0; k < n; ++k) {
                const int s = sHi - 1 - k, p = s % T, p1 = (s + 1) % T, it = itBase + k, st = it % NST;
                mbar_wait(&sm.mbar[st], (uint32_t)((it / NST) & 1));
                // contribution of block s+1 (x written by this warp one iteration ago): rows 2q, 2q+1 of slot p1, column c
                cplx acc = mk(0.0, 0.0);
                if (T > 1) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int r = p1 * TS + 2 * q + e;
                        cfma(acc, mk(sm.stage[st][0][c >> 2][r][c & 3], sm.stage[st][1][c >> 2][r][c & 3]), sm.y[r]);
                    }
                }
                bar_sync(SB_F0 + (k & 1), NTHR);                         // far(s): the rows of blocks s+2 .. s+T-1
                for (int w = q; w < NTHR / 32 - 2; w += 4) acc += sm.part[k & 1][w][c];
                reduce_rows_in_warp(acc);                                 // d_c in every lane with this c
                const int ii = lane >> 2, tt = lane & 3;     // lane handles columns 2tt, 2tt+1 of row ii
                const cplx d0 = mk(__shfl_sync(0xffffffffu, acc.x, 2 * tt), __shfl_sync(0xffffffffu, acc.y, 2 * tt));
                const cplx d1 = mk(__shfl_sync(0xffffffffu, acc.x, 2 * tt + 1), __shfl_sync(0xffffffffu, acc.y, 2 * tt + 1));
                cplx xv = sm.az[st][ii * 8 + 2 * tt] * d0 + sm.az[st][ii * 8 + 2 * tt + 1] * d1;
                reduce_quad(xv);
                const cplx xo = sm.az[st][64 + ii] - xv;
                if (tt == 0) sm.y[p * TS + ii] = xo;
                __syncwarp();                                             // x_s is read by other lanes of this warp in the next iteration
                if (k + 2 < n) bar_arrive(SB_X0 + (k & 1), NTHR);        // x_s published: far(s-2) may start (every arrive has its wait)
            }
        } else {
            // far warps (1..6): partial dot products; service warp (7): lanes 0..7 store x of step k-2, lane 8 refills its TMA
            // stage (z_s travels with the panel).  One loop for both, see the forward sweep.
            const bool service = warp == NTHR / 32 - 1;
            const int ft = tid - 32, c = ft & 7, rg = ft >> 3;
            constexpr int NRG = (NTHR - 64) / 8;          // row groups of the far compute warps
            for (int k = 0; k < n; ++k) {
                if (k >= 2) bar_sync(SB_X0 + (k & 1), NTHR);              // x of step k-2 published (the first two jobs need no new x)
                const int s = sHi - 1 - k, p = s % T, p1 = (s + 1) % T, it = itBase + k, st = it % NST;
                if (service) {
                    if (k >= 2 && lane < 8) local_store(L, job.x, (s + 2) * TS + lane, sm.y[((s + 2) % T) * TS + lane]);
                    if (k >= 2 && lane == 8 && k - 2 + NST < n) issue(s + 2 - NST, it - 2 + NST, -1);
                } else {
                    mbar_wait(&sm.mbar[st], (uint32_t)((it / NST) & 1));
                    cplx a0 = mk(0.0, 0.0), a1 = mk(0.0, 0.0);
#pragma unroll
                    for (int r_b = 0; r_b < R; r_b += 2 * NRG) {
                        const int r0 = r_b + rg, r1 = r0 + NRG;
                        if (r0 < R && (r0 >> 3) != p && (r0 >> 3) != p1)
                            cfma(a0, mk(sm.stage[st][0][c >> 2][r0][c & 3], sm.stage[st][1][c >> 2][r0][c & 3]), sm.y[r0]);
                        if (r1 < R && (r1 >> 3) != p && (r1 >> 3) != p1)
                            cfma(a1, mk(sm.stage[st][0][c >> 2][r1][c & 3], sm.stage[st][1][c >> 2][r1][c & 3]), sm.y[r1]);
                    }
                    cplx acc = a0 + a1;
                    reduce_rows_in_warp(acc);
                    if (lane < 8) sm.part[k & 1][warp - 1][lane] = acc;
                }
                bar_arrive(SB_F0 + (k & 1), NTHR);
            }
        }
        cta_sync();
        // the solutions of the last two steps (the service warp stores x of step k-2 while step k runs)
        if (warp == NTHR / 32 - 1 && lane < 8)
            for (int kk = (n >= 2 ? n - 2 : 0); kk < n; ++kk) {
                const int sx = sHi - 1 - kk;
                local_store(L, job.x, sx * TS + lane, sm.y[(sx % T) * TS + lane]);
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
    fence_proxy_async_all();      // zbuf was written by this CTA during the forward sweep and is read back through TMA
    __threadfence();              // (earlier launches are ordered by the stream)
    cta_sync();
    for (int i_b = 0; i_b < R; i_b += NTHR) if (const int i = i_b + tid; i < R)
        sm.y[i] = (mode == FM_BACK || mode == SM_BACKZ_OWN) ? xsep[rel(i >> 3, L.sOwn) * TS + (i & 7)] : mk(0.0, 0.0);
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
