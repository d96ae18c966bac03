// Large-bandwidth path of the batched complex-symmetric block-LDL^T factorisation (half-bandwidth 105 .. 320,
// e.g. the 800 x 300-cell mesh of BASELINE.json configs[3], b = 299).
//
// Same algorithm and the SAME stored factor as band_factor.cuh — 8x8 pivot blocks, per macro-step
// { raw panel image, A11^{-1}, z } — so the solve kernel (band_solve.cuh) serves both paths.  What changes is where
// the sliding window lives: T x T tiles no longer fit one SM's register file, so the window is a circular R x R
// array in global memory (L2-resident: 1.6 MB per system at b = 299) and the elimination advances kBigNBK = 4
// macro-steps (32 columns) per pair of stream-ordered launches:
//   bigband_panel_kernel   kBigSplit CTAs per system: each loads the 32 pivot rows of the panel plus its share of the other
//                          window rows into shared memory, eliminates the four 8x8 pivot blocks one after the other
//                          (same in-register Gauss-Jordan as the register kernel; redundantly in every CTA, so the CTAs
//                          never talk), forms M' = -raw A11^{-1} and the in-panel updates for its row blocks with
//                          DMMA.8x8x4 (one warp per 8-row block), forward-eliminates the fused right-hand side and
//                          writes its rows of the four factor panels;
//   bigband_update_kernel  many CTAs per system: rank-32 update of the trailing window with FP64 tensor-core MMAs
//                          (DMMA.8x8x4), one warp per 8x8 tile, operands straight from L2/L1; four extra CTAs
//                          refill the recycled window rows from the stencil (assembly stays fused, the matrix is
//                          never stored).
// No cross-CTA synchronisation inside a kernel.  Replaces factorMUMPS / lu for the systems of mt2DTE.jl:47-55,
// mt2DTM.jl:46-54 when the mesh is too wide for the register-window kernel.
#pragma once
#include "band_factor.cuh"

namespace hmcmt {

constexpr int kBigNBK = 4;          // 8-column blocks per panel launch
constexpr int kBigPanStride = 9;    // complex entries per shared-memory panel row (8 + 1 pad: conflict-free 16-byte accesses)
constexpr int kBigMaxT = 44;        // window of 44 tiles: b <= 8 * (44 - kBigNBK) = 320
constexpr int kBigSplit = 8;        // CTAs per system in the panel kernel

__host__ __device__ constexpr int big_T_for(int b) { return ((b + 7) / 8 + kBigNBK + 3) / 4 * 4; }
// per-system workspace (complex entries): window R*R | M' [NBK][R][8] | raw [NBK][R][8] | rhs window [2][R] (by panel parity)
__host__ __device__ constexpr size_t big_work_entries(int T) {
    return (size_t)(TS * T) * (TS * T) + 2 * (size_t)kBigNBK * (TS * T) * 8 + 2 * (size_t)(TS * T);
}
// local rows of one panel CTA: the 4 pivot blocks + its share of the other window blocks (one warp per 8-row block)
__host__ __device__ constexpr int big_panel_local_rows(int T) {
    return TS * (kBigNBK + (T - kBigNBK + kBigSplit - 1) / kBigSplit);
}
__host__ __device__ constexpr size_t big_panel_smem_bytes(int T) {
    return ((size_t)kBigNBK * big_panel_local_rows(T) * kBigPanStride + (size_t)big_panel_local_rows(T) + 64 + 8 +
            (size_t)big_panel_local_rows(T) * kBigPanStride) * sizeof(cplx);
}

struct BigView {
    cplx *win, *mscr, *rscr, *ywin;
    __device__ __forceinline__ BigView(cplx* base, int R) {
        win = base;
        mscr = win + (size_t)R * R;
        rscr = mscr + (size_t)kBigNBK * R * 8;
        ywin = rscr + (size_t)kBigNBK * R * 8;
    }
};

// global block held by window slot `slot` when the window covers blocks s0 .. s0+T-1
__device__ __forceinline__ int big_block_of_slot(int slot, int s0, int T) {
    int a = slot - s0 % T;
    if (a < 0) a += T;
    return s0 + a;
}

// Pristine matrix entries for the 8 window rows of block `beta` (slot beta % T); the window covers blocks s0w .. s0w+T-1.
// Cell convention: entry (row block bi >= column block bj) lives at win[(slot(bi)*8 + i) * R + slot(bj)*8 + j].
__device__ __forceinline__ void big_fill_rows(const EntryProvider& prov, cplx* win, int T, int s0w, int beta, int tid, int nthr) {
    const int R = TS * T, slot = beta % T;
    for (int e = tid; e < TS * R; e += nthr) {
        const int i = e / R, col = e - i * R;
        const int bj = big_block_of_slot(col >> 3, s0w, T);
        if (bj > beta) continue;
        const int gi = beta * TS + i, gj = bj * TS + (col & 7);
        win[(size_t)(slot * TS + i) * R + col] = prov.get(max(gi, gj), min(gi, gj));
    }
}

// grid (T, nsys), 256 threads: initial window (blocks 0 .. T-1) and rhs window
__global__ void __launch_bounds__(256)
bigband_init_kernel(const BandSys* __restrict__ systems, BandDom dom, int T) {
    const BandSys sys = systems[blockIdx.y];
    const LocalDom L = LocalDom::make(dom, 0);
    EntryProvider prov{sys.dr, sys.dm, sys.e1, sys.e2, sys.band, sys.omega, dom.b, L};
    const int R = TS * T;
    BigView v(sys.big, R);
    big_fill_rows(prov, v.win, T, 0, (int)blockIdx.x, threadIdx.x, blockDim.x);
    if (blockIdx.x == 0)
        for (int r = threadIdx.x; r < R; r += blockDim.x)
            v.ywin[r] = (sys.rhs && r < L.nLoc) ? prov.rhs_at(sys.rhs, r) : mk(0.0, 0.0);
}

// grid (kBigSplit, nsys).  The 32 pivot rows of the panel (its 32x32 diagonal block) are factored redundantly by every CTA
// of a system — that is the serial part, four 8x8 inversions — and the remaining window row blocks are divided among the CTAs:
// given the diagonal block's factors the row blocks are independent, so there is no communication between the CTAs.
// Local row blocks: 0..3 = pivot blocks, then this CTA's share of the other window blocks.  One warp per local 8-row block:
// M' = raw (-A11^{-1}) and the in-panel updates are 8x8x8 complex products on the FP64 tensor cores (DMMA.8x8x4, the
// register kernel's formulation).  CTA 0 alone writes what belongs to the pivot rows.
// The rhs window is double-buffered by panel parity (every CTA reads all pivot rows of it while CTA 0 rewrites them).
// block = 32 * big_panel_local_rows(T) / 8 threads, dynamic smem = big_panel_smem_bytes(T)
__global__ void __launch_bounds__(4 * big_panel_local_rows(kBigMaxT))
bigband_panel_kernel(const BandSys* __restrict__ systems, BandDom dom, int T, int k) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int R = TS * T;
    constexpr int ND = kBigNBK * TS;                                      // pivot rows
    const int chunkBlocks = (T - kBigNBK + kBigSplit - 1) / kBigSplit;   // other row blocks per CTA
    const int NL = big_panel_local_rows(T);                               // local rows = ND + 8 * chunkBlocks
    cplx* const pan = reinterpret_cast<cplx*>(smem_raw);                  // [NBK][NL][kBigPanStride]
    cplx* const ysh = pan + (size_t)kBigNBK * NL * kBigPanStride;         // [NL]
    cplx* const ainv = ysh + NL;                                          // [64] row-major A11^{-1}
    cplx* const zsh = ainv + 64;                                          // [8]
    cplx* const mwAll = zsh + 8;                                          // [NL/8][8][kBigPanStride]: M' of each warp's row block

    const BandSys sys = systems[blockIdx.y];
    const LocalDom L = LocalDom::make(dom, 0);
    EntryProvider prov{sys.dr, sys.dm, sys.e1, sys.e2, sys.band, sys.omega, dom.b, L};
    BigView v(sys.big, R);
    const cplx* const yin = v.ywin + (size_t)(k & 1) * R;
    cplx* const yout = v.ywin + (size_t)((k + 1) & 1) * R;
    const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int s0 = k * kBigNBK, q0 = s0 % T;                              // the pivot slots are q0 .. q0+3 (T % 4 == 0: no wrap)
    const int nsub = min(kBigNBK, L.sTot - s0);
    const bool first = blockIdx.x == 0;
    // window slot of a local row block (-1: none): pivot blocks first, then this CTA's share of the others in window order
    auto slot_of = [&](int lb) {
        if (lb < kBigNBK) return q0 + lb;
        const int ob = (int)blockIdx.x * chunkBlocks + (lb - kBigNBK);   // index among the T - NBK non-pivot blocks
        if (lb - kBigNBK >= chunkBlocks || ob >= T - kBigNBK) return -1;
        return ob < q0 ? ob : ob + kBigNBK;
    };
    auto P = [&](int c, int row, int col) -> cplx& { return pan[((size_t)c * NL + row) * kBigPanStride + col]; };

    for (int idx = tid; idx < 2 * NL; idx += blockDim.x) {
        const int lr = idx >> 1, h = idx & 1, slot = slot_of(lr >> 3);
        const int r = slot < 0 ? -1 : slot * TS + (lr & 7);
        const int beta = slot < 0 ? -1 : big_block_of_slot(slot, s0, T);
        for (int c = 0; c < kBigNBK; ++c) {
            const bool live = r >= 0 && c < nsub && beta >= s0 + c;
#pragma unroll
            for (int j = 0; j < 4; ++j) P(c, lr, 4 * h + j) = live ? v.win[(size_t)r * R + (q0 + c) * TS + 4 * h + j] : mk(0.0, 0.0);
        }
        if (h == 0) ysh[lr] = r >= 0 ? yin[r] : mk(0.0, 0.0);
    }
    __syncthreads();

    // this warp's row block (warp-uniform)
    const int lr0 = warp * TS, slot = slot_of(warp);
    const int beta = slot < 0 ? -1 : big_block_of_slot(slot, s0, T);
    const bool owner = slot >= 0 && (warp >= kBigNBK || first);           // writes this row block's results
    const int rg = slot * TS + g;                                         // window row of lane group g
    cplx* const mw = mwAll + (size_t)warp * TS * kBigPanStride;
    for (int c = 0; c < nsub; ++c) {
        if (warp == 0) {
            const int i = g;
            cplx a0 = P(c, c * TS + i, 2 * t), a1 = P(c, c * TS + i, 2 * t + 1);
            bool bad = false;
            gj_invert8<true>(a0, a1, bad, i, t);
            bad = __any_sync(0xffffffffu, bad);
            if (bad && lane == 0 && first && sys.status) *sys.status = kErrSingular;
            ainv[i * 8 + 2 * t] = a0;
            ainv[i * 8 + 2 * t + 1] = a1;
            cplx acc = a0 * ysh[c * TS + 2 * t] + a1 * ysh[c * TS + 2 * t + 1];
#pragma unroll
            for (int off = 1; off <= 2; off <<= 1) {
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
                acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
            }
            if (!sys.rhs) acc = mk(0.0, 0.0);
            if (t == 0) zsh[i] = acc;
            if (first) {
                cplx* dz = sys.ainvz[0] + (size_t)(s0 + c) * AZ;
                dz[i * 8 + 2 * t] = a0;
                dz[i * 8 + 2 * t + 1] = a1;
                if (t == 0) dz[64 + i] = acc;
            }
        }
        __syncthreads();
        const bool below = beta > s0 + c, keep = beta >= s0 + c;
        // M'[X] = raw_c[X] (-A11^{-1}): lane (g,t) ends up with entries (g, 2t) and (g, 2t+1)
        cplx m0 = mk(0.0, 0.0), m1 = mk(0.0, 0.0);
        if (below) {
            double mre[2] = {0.0, 0.0}, mim[2] = {0.0, 0.0}, t1[2] = {0.0, 0.0}, t2[2] = {0.0, 0.0};
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const cplx a = P(c, lr0 + g, 4 * kk + t);
                const cplx bq = ainv[(4 * kk + t) * 8 + g];
                dmma884(mre, a.x, -bq.x);
                dmma884(mim, a.x, -bq.y);
                dmma884(t1, -a.y, -bq.y);
                dmma884(t2, a.y, -bq.x);
            }
            m0 = mk(mre[0] + t1[0], mim[0] + t2[0]);
            m1 = mk(mre[1] + t1[1], mim[1] + t2[1]);
            // fused forward elimination of the rhs: y_X -= raw_c[X] z
            if (t == 0 && sys.rhs) {
                cplx acc = ysh[lr0 + g];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) cfma(acc, -P(c, lr0 + g, kk), zsh[kk]);
                ysh[lr0 + g] = acc;
            }
        }
        mw[g * kBigPanStride + 2 * t] = m0;
        mw[g * kBigPanStride + 2 * t + 1] = m1;
        // operands of the trailing update (plain [c][r][8] layout, L2-resident)
        if (owner) {
            v.mscr[((size_t)c * R + rg) * 8 + 2 * t] = m0;
            v.mscr[((size_t)c * R + rg) * 8 + 2 * t + 1] = m1;
        }
        __syncwarp();
        if (below) {
            double are[2], aim[2];
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const cplx a = mw[g * kBigPanStride + 4 * kk + t];
                are[kk] = a.x; aim[kk] = a.y;
            }
            for (int c2 = c + 1; c2 < nsub; ++c2) {
                if (beta < s0 + c2) continue;                            // (warp-uniform)
                const cplx c0 = P(c2, lr0 + g, 2 * t), c1 = P(c2, lr0 + g, 2 * t + 1);
                double cre[2] = {c0.x, c1.x}, cim[2] = {c0.y, c1.y}, t1[2] = {0.0, 0.0}, t2[2] = {0.0, 0.0};
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const cplx bq = P(c, c2 * TS + g, 4 * kk + t);     // raw_c restricted to the rows of pivot block c2
                    dmma884(cre, are[kk], bq.x);
                    dmma884(cim, are[kk], bq.y);
                    dmma884(t1, -aim[kk], bq.y);
                    dmma884(t2, aim[kk], bq.x);
                }
                P(c2, lr0 + g, 2 * t) = mk(cre[0] + t1[0], cim[0] + t2[0]);
                P(c2, lr0 + g, 2 * t + 1) = mk(cre[1] + t1[1], cim[1] + t2[1]);
            }
        }
        if (owner) {
            const cplx e0 = P(c, lr0 + g, 2 * t), e1 = P(c, lr0 + g, 2 * t + 1);
            v.rscr[((size_t)c * R + rg) * 8 + 2 * t] = below ? e0 : mk(0.0, 0.0);
            v.rscr[((size_t)c * R + rg) * 8 + 2 * t + 1] = below ? e1 : mk(0.0, 0.0);
            // factor panel of macro-step s0+c in the solve kernels' image layout [re/im][kk][R][4]; rows of blocks already
            // eliminated in this launch are outside the band of this panel: zero
            double* img = sys.panels[0] + (size_t)(s0 + c) * panel_doubles(T);
            const int kkI = t >> 1, tt = 2 * (t & 1);
            *reinterpret_cast<double2*>(img + ((size_t)(0 * 2 + kkI) * R + rg) * 4 + tt) = keep ? make_double2(e0.x, e1.x) : make_double2(0.0, 0.0);
            *reinterpret_cast<double2*>(img + ((size_t)(1 * 2 + kkI) * R + rg) * 4 + tt) = keep ? make_double2(e0.y, e1.y) : make_double2(0.0, 0.0);
        }
        __syncthreads();
    }
    // rhs window of the next panel: the slots of the eliminated blocks now hold blocks +T
    if (owner && t == 0) {
        cplx yv = ysh[lr0 + g];
        if (beta < s0 + nsub) {
            const int gr = (beta + T) * TS + g;
            yv = (sys.rhs && gr < L.nLoc) ? prov.rhs_at(sys.rhs, gr) : mk(0.0, 0.0);
        }
        yout[rg] = yv;
    }
}

// grid (nXp * nYq + kBigNBK, nsys), 256 threads.  CTAs x < nXp*nYq: a 2 x 4 block of 8x8 tiles of the trailing window, one
// tile per warp (the two M' row blocks and four raw row blocks of a CTA are shared through L1); CTAs beyond refill one
// recycled 8-row block from the stencil.  nXp = ceil(npos/2), nYq = ceil(npos/4), npos = T - kBigNBK window positions.
__global__ void __launch_bounds__(256)
bigband_update_kernel(const BandSys* __restrict__ systems, BandDom dom, int T, int k, int nYq, int nBlocks) {
    const BandSys sys = systems[blockIdx.y];
    const int R = TS * T;
    BigView v(sys.big, R);
    const int s0 = k * kBigNBK, w0 = s0 + kBigNBK;      // first block of the trailing window
    if ((int)blockIdx.x >= nBlocks) {
        const LocalDom L = LocalDom::make(dom, 0);
        EntryProvider prov{sys.dr, sys.dm, sys.e1, sys.e2, sys.band, sys.omega, dom.b, L};
        big_fill_rows(prov, v.win, T, w0, s0 + ((int)blockIdx.x - nBlocks) + T, threadIdx.x, blockDim.x);
        return;
    }
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int warp = __shfl_sync(0xffffffffu, (int)threadIdx.x >> 5, 0);
    const int xb = (int)blockIdx.x / nYq, yb = (int)blockIdx.x - xb * nYq;
    const int aX = 2 * xb + (warp >> 2), aY = 4 * yb + (warp & 3);
    if (aX >= T - kBigNBK || aY > aX) return;
    const int slotX = (w0 + aX) % T, slotY = (w0 + aY) % T;
    double are[kBigNBK][2], aim[kBigNBK][2], bre[kBigNBK][2], bim[kBigNBK][2];
#pragma unroll
    for (int c = 0; c < kBigNBK; ++c)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            const cplx a = __ldg(&v.mscr[((size_t)c * R + slotX * TS + g) * 8 + 4 * kk + t]);
            const cplx bv = __ldg(&v.rscr[((size_t)c * R + slotY * TS + g) * 8 + 4 * kk + t]);
            are[c][kk] = a.x; aim[c][kk] = a.y;
            bre[c][kk] = bv.x; bim[c][kk] = bv.y;
        }
    cplx* cp = v.win + (size_t)(slotX * TS + g) * R + slotY * TS + 2 * t;
    const cplx c0 = cp[0], c1 = cp[1];
    double cre[2] = {c0.x, c1.x}, cim[2] = {c0.y, c1.y}, t1[2] = {0.0, 0.0}, t2[2] = {0.0, 0.0};
#pragma unroll
    for (int c = 0; c < kBigNBK; ++c)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            dmma884(cre, are[c][kk], bre[c][kk]);
            dmma884(cim, are[c][kk], bim[c][kk]);
            dmma884(t1, -aim[c][kk], bim[c][kk]);
            dmma884(t2, aim[c][kk], bre[c][kk]);
        }
    cp[0] = mk(cre[0] + t1[0], cim[0] + t2[0]);
    cp[1] = mk(cre[1] + t1[1], cim[1] + t2[1]);
}

}  // namespace hmcmt
