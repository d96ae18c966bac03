// Numeric phase of the nested-dissection multifrontal solver: batched dense partial factorisations of the fronts on the
// FP64 tensor cores (DMMA.8x8x4) and the level-by-level triangular solves.  Replaces `factorMUMPS(Aii,1)` / `applyMUMPS`
// (mt2DTE.jl:50-53, compJacTMatVec.jl:220-224, MUMPSfuncs.jl:24-39,75-132) for systems whose half-bandwidth exceeds the
// register-window kernel (band_factor.cuh) and for arbitrary symmetric matrices handed to the Level-1 shim.
//
// Front arithmetic (pivot-free, complex symmetric, no conjugation; prototype tools/proto/mf_proto.py), front = [pivots | update]:
//     G = F11^{-1},   M = F21 G,   U = F22 - M F21^T          (the block "sweep" of the pivots; 8x8 pivot blocks inverted in
//                                                               registers by gj_invert8)
//     forward  w2 -= M w1 ;   backward  x1 = G w1 - M^T x2       — no triangular solves: every step is a dense product.
//   * small fronts (fp <= fSmall): ONE CTA assembles the front in shared memory (original entries + extend-add of the
//     children's update matrices), sweeps it tile by tile and writes G, M (factor arena) and U (update arena);
//   * large fronts live in global memory: assembly kernels, then per chunk of <= 96 pivots: mf_inv_kernel (G of the diagonal
//     block, shared memory), mf_gemm_kernel (M = F21 G), mf_gemm_kernel (trailing update U -= M F21^T) — cp.async-staged
//     64x64x16 tiles, 8 warps, DMMA.
//
// Matrix storage ("k-grouped"): element (i,j) of a matrix with ld rows sits at ((j/4*2 + plane)*ld + i)*4 + j%4 doubles
// (plane 0 = real, 1 = imaginary).  A block of rows of four consecutive columns is contiguous — exactly the DMMA A/B operand
// image (lane (g,t) reads [row g][k t]) — so operand tiles are copied global -> shared with 16-byte cp.async and no transposition,
// and C fragments (lane holds [g][2t], [g][2t+1]) are read / written as 16-byte pairs.
#pragma once
#include "band_factor.cuh"
#include "mf_solver.cuh"

namespace hmcmt {
namespace mf {

__host__ __device__ __forceinline__ size_t kg_off(int ld, int i, int j, int plane) {
    return ((size_t)((j >> 2) * 2 + plane) * ld + i) * 4 + (j & 3);
}

// device view of the symbolic tables + per-batch arenas
struct Tables {
    const Front* fronts;
    const int* rows;
    const int* rel;
    const int* children;
    const OrigEntry* orig;
    const Chunk* chunks;
    const int* pos2orig;
    double* fac;            // [nsys][facStride]
    double* arena[2];       // [nsys][arenaStride[p]]
    int64_t facStride, arenaStride[2];
    const cplx* vals;       // [nsys][valStride]
    int64_t valStride;
    int* status;            // [nsys]
    unsigned long long* prof;   // optional (HMCMT_MF_PROF=1): per-phase cycle totals of mf_small_kernel, thread 0 of every CTA
};
#define MF_PROF_MARK(slot) do { if (tb.prof && threadIdx.x == 0) { const long long _t = clock64(); atomicAdd(tb.prof + _pc * 8 + (slot), (unsigned long long)(_t - _t0)); _t0 = _t; } } while (0)

// ------------------------------------------------------------------------------------------------------------------------
// block sweep of the first npb pivot blocks of an nb x nb tile matrix held in shared memory (lower-triangle tiles, 128 doubles
// each: [plane][8][8]; of a diagonal tile only the lower triangle is meaningful: it is always read mirrored).  On return: pivot x pivot tiles hold -G, rest x pivot tiles hold M, rest x rest U.
template <int NW>
__device__ __forceinline__ void mf_sweep(double* __restrict__ tiles, double* __restrict__ raw, double* __restrict__ mm,
                                         double* __restrict__ nainv, int* __restrict__ fail, const int nb, const int npb,
                                         unsigned long long* __restrict__ prof = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31;
    long long _s0 = prof ? clock64() : 0;
#define MF_SWEEP_MARK(slot) do { if (prof && tid == 0) { const long long _t = clock64(); atomicAdd(prof + (slot), (unsigned long long)(_t - _s0)); _s0 = _t; } } while (0)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = lane >> 2, t = lane & 3;
    const int R = nb * 8;
    const int nT = nb * (nb + 1) / 2;
    auto tileP = [&](int I, int J) { return tiles + (size_t)(I * (I + 1) / 2 + J) * 128; };
    auto opnd = [&](double* base, int pl, int kk, int row) { return base + ((size_t)(pl * 2 + kk) * R + row) * 4; };
    // lane roles of the tile <-> operand-panel copies: plane, row, k-half
    const int cpl = lane >> 4, crow = (lane >> 1) & 7, ckk = lane & 1;
    for (int kb = 0; kb < npb; ++kb) {
        // (a) column block kb of the symmetric matrix as an operand panel: raw_I = A[I][kb]  (one tile per warp and trip)
        for (int I = warp; I < nb; I += NW) {
            double2 v01, v23;
            if (I > kb) {
                const double* src = tileP(I, kb) + cpl * 64 + crow * 8 + ckk * 4;
                v01 = *reinterpret_cast<const double2*>(src);
                v23 = *reinterpret_cast<const double2*>(src + 2);
            } else if (I == kb) {          // diagonal tile: only its lower triangle is maintained -> mirrored read
                const double* D = tileP(kb, kb) + cpl * 64;
                const int c0 = ckk * 4;
                auto sym = [&](int r, int c) { return r >= c ? D[r * 8 + c] : D[c * 8 + r]; };
                v01 = make_double2(sym(crow, c0), sym(crow, c0 + 1));
                v23 = make_double2(sym(crow, c0 + 2), sym(crow, c0 + 3));
            } else {
                const double* src = tileP(kb, I) + cpl * 64 + (ckk * 4) * 8 + crow;
                v01 = make_double2(src[0], src[8]);
                v23 = make_double2(src[16], src[24]);
            }
            double* dst = opnd(raw, cpl, ckk, I * 8 + crow);
            *reinterpret_cast<double2*>(dst) = v01;
            *reinterpret_cast<double2*>(dst + 2) = v23;
        }
        // (b) P = A[kb][kb]^{-1} in registers (warp 0 reads the diagonal tile itself: no barrier needed before);
        //     publish -P as the B operand and as the new diagonal tile
        if (warp == 0) {
            __syncwarp();
            double* D = tileP(kb, kb);
            auto symre = [&](int r, int c) { return r >= c ? D[r * 8 + c] : D[c * 8 + r]; };
            auto symim = [&](int r, int c) { return r >= c ? D[64 + r * 8 + c] : D[64 + c * 8 + r]; };
            cplx a0 = mk(symre(g, 2 * t), symim(g, 2 * t)), a1 = mk(symre(g, 2 * t + 1), symim(g, 2 * t + 1));
            bool bad = false;
            gj_invert8<true>(a0, a1, bad, g, t);
            if (__any_sync(0xffffffffu, bad) && lane == 0) *fail = 1;
            const int j0 = 2 * t;
            *reinterpret_cast<double2*>(nainv + ((0 * 2 + (j0 >> 2)) * 8 + g) * 4 + (j0 & 3)) = make_double2(-a0.x, -a1.x);
            *reinterpret_cast<double2*>(nainv + ((1 * 2 + (j0 >> 2)) * 8 + g) * 4 + (j0 & 3)) = make_double2(-a0.y, -a1.y);
        }
        MF_SWEEP_MARK(0);
        __syncthreads();
        MF_SWEEP_MARK(1);
        if (warp == 0) {       // the raw copy of the diagonal tile is complete: the tile itself may now take -P
            double* D = tileP(kb, kb);
            const int j0 = 2 * t;
            const double2 pre = *reinterpret_cast<const double2*>(nainv + ((0 * 2 + (j0 >> 2)) * 8 + g) * 4 + (j0 & 3));
            const double2 pim = *reinterpret_cast<const double2*>(nainv + ((1 * 2 + (j0 >> 2)) * 8 + g) * 4 + (j0 & 3));
            *reinterpret_cast<double2*>(D + g * 8 + 2 * t) = pre;
            *reinterpret_cast<double2*>(D + 64 + g * 8 + 2 * t) = pim;
        }
        // (c) m_I = raw_I (-P) for every block row I != kb
        for (int I = warp; I < nb; I += NW) {
            if (I == kb) continue;
            double mre[2] = {0.0, 0.0}, mim[2] = {0.0, 0.0}, mr2[2] = {0.0, 0.0}, mi2[2] = {0.0, 0.0};
            const int r = I * 8 + g;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const double are = opnd(raw, 0, kk, r)[t], aim = opnd(raw, 1, kk, r)[t];
                const double bre = nainv[((0 * 2 + kk) * 8 + g) * 4 + t], bim = nainv[((1 * 2 + kk) * 8 + g) * 4 + t];
                dmma884(mre, are, bre);
                dmma884(mim, are, bim);
                dmma884(mr2, -aim, bim);
                dmma884(mi2, aim, bre);
            }
            mre[0] += mr2[0]; mre[1] += mr2[1]; mim[0] += mi2[0]; mim[1] += mi2[1];
            *reinterpret_cast<double2*>(opnd(mm, 0, t >> 1, r) + (t & 1) * 2) = make_double2(mre[0], mre[1]);
            *reinterpret_cast<double2*>(opnd(mm, 1, t >> 1, r) + (t & 1) * 2) = make_double2(mim[0], mim[1]);
        }
        __syncthreads();
        MF_SWEEP_MARK(2);
        // (d) A[I][J] += m_I raw_J^T for I,J != kb ;  (e) column kb <- raw P = -m.   Lower tiles dealt round-robin to the warps;
        // trailing updates are issued two tiles at a time (16 independent DMMAs in flight: their latency is ~138 cycles).
        auto upd_load = [&](double* T, double (&cre)[2], double (&cim)[2]) {
            const double2 cr = *reinterpret_cast<const double2*>(T + g * 8 + 2 * t);
            const double2 ci = *reinterpret_cast<const double2*>(T + 64 + g * 8 + 2 * t);
            cre[0] = cr.x; cre[1] = cr.y; cim[0] = ci.x; cim[1] = ci.y;
        };
        auto upd_mma = [&](int I_, int J_, double (&cre)[2], double (&cim)[2], double (&t1)[2], double (&t2)[2]) {
            const int ra = I_ * 8 + g, rb = J_ * 8 + g;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const double are = opnd(mm, 0, kk, ra)[t], aim = opnd(mm, 1, kk, ra)[t];
                const double bre = opnd(raw, 0, kk, rb)[t], bim = opnd(raw, 1, kk, rb)[t];
                dmma884(cre, are, bre);
                dmma884(cim, are, bim);
                dmma884(t1, -aim, bim);
                dmma884(t2, aim, bre);
            }
        };
        auto upd_store = [&](double* T, const double (&cre)[2], const double (&cim)[2], const double (&t1)[2], const double (&t2)[2]) {
            *reinterpret_cast<double2*>(T + g * 8 + 2 * t) = make_double2(cre[0] + t1[0], cre[1] + t1[1]);
            *reinterpret_cast<double2*>(T + 64 + g * 8 + 2 * t) = make_double2(cim[0] + t2[0], cim[1] + t2[1]);
        };
        int I = 0, J = warp;
        while (J > I) { J -= I + 1; ++I; }
        int pI = -1, pJ = -1;          // pending trailing-update tile (waiting for a partner)
        for (int L = warp; L < nT; L += NW) {
            double* T = tileP(I, J);
            if (I == kb && J == kb) {
            } else if (J == kb) {              // I > kb
                const double2 a = *reinterpret_cast<const double2*>(opnd(mm, 0, t >> 1, I * 8 + g) + (t & 1) * 2);
                const double2 b = *reinterpret_cast<const double2*>(opnd(mm, 1, t >> 1, I * 8 + g) + (t & 1) * 2);
                *reinterpret_cast<double2*>(T + g * 8 + 2 * t) = make_double2(-a.x, -a.y);
                *reinterpret_cast<double2*>(T + 64 + g * 8 + 2 * t) = make_double2(-b.x, -b.y);
            } else if (I == kb) {              // J < kb: transposed
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    T[g * 8 + 2 * t + e] = -opnd(mm, 0, g >> 2, J * 8 + 2 * t + e)[g & 3];
                    T[64 + g * 8 + 2 * t + e] = -opnd(mm, 1, g >> 2, J * 8 + 2 * t + e)[g & 3];
                }
            } else if (pI < 0) {
                pI = I; pJ = J;
            } else {
                double* T0 = tileP(pI, pJ);
                double c0r[2], c0i[2], c1r[2], c1i[2], a1[2] = {0.0, 0.0}, a2[2] = {0.0, 0.0}, b1[2] = {0.0, 0.0}, b2[2] = {0.0, 0.0};
                upd_load(T0, c0r, c0i);
                upd_load(T, c1r, c1i);
                upd_mma(pI, pJ, c0r, c0i, a1, a2);
                upd_mma(I, J, c1r, c1i, b1, b2);
                upd_store(T0, c0r, c0i, a1, a2);
                upd_store(T, c1r, c1i, b1, b2);
                pI = -1;
            }
            J += NW;
            while (J > I) { J -= I + 1; ++I; }
        }
        if (pI >= 0) {
            double* T0 = tileP(pI, pJ);
            double c0r[2], c0i[2], a1[2] = {0.0, 0.0}, a2[2] = {0.0, 0.0};
            upd_load(T0, c0r, c0i);
            upd_mma(pI, pJ, c0r, c0i, a1, a2);
            upd_store(T0, c0r, c0i, a1, a2);
        }
        __syncthreads();
        MF_SWEEP_MARK(3);
    }
#undef MF_SWEEP_MARK
}

__host__ __device__ inline size_t mf_sweep_smem_bytes(int nb) {
    return ((size_t)nb * (nb + 1) / 2 * 128 + 2 * 16 * (size_t)nb * 8 + 128) * sizeof(double) + 16;
}

// element (a,b) of the symmetric tile matrix, a,b local indices
__device__ __forceinline__ cplx mf_tile_get(const double* tiles, int a, int b) {
    if (a < b) { const int x = a; a = b; b = x; }
    const int I = a >> 3, J = b >> 3;
    const double* T = tiles + (size_t)(I * (I + 1) / 2 + J) * 128 + (a & 7) * 8 + (b & 7);
    return mk(T[0], T[64]);
}

// one 8x8 tile (shared memory, [plane][8][8]) -> k-grouped global matrix with ld rows at (row0, col0); executed by one warp:
// lane = (chunk, row): chunk = (plane, column group of four), 32 contiguous bytes per lane, 256 per eight lanes.
// TRANSPOSED: the tile holds the transposed block.  sign: +1 / -1.
// MODE 0: as stored, 1: transposed, 2: diagonal tile (lower triangle mirrored)
template <int MODE>
__device__ __forceinline__ void mf_store_tile(const double* __restrict__ T, double* __restrict__ dst, int ld, int row0, int col0,
                                              double sign, int lane) {
    const int pl = lane >> 4, cg = (lane >> 3) & 1, row = lane & 7;
    double2 v01, v23;
    if (MODE == 0) {
        const double* src = T + pl * 64 + row * 8 + cg * 4;
        v01 = *reinterpret_cast<const double2*>(src);
        v23 = *reinterpret_cast<const double2*>(src + 2);
    } else if (MODE == 2) {
        const double* D = T + pl * 64;
        const int c0 = cg * 4;
        auto sym = [&](int r, int c) { return r >= c ? D[r * 8 + c] : D[c * 8 + r]; };
        v01 = make_double2(sym(row, c0), sym(row, c0 + 1));
        v23 = make_double2(sym(row, c0 + 2), sym(row, c0 + 3));
    } else {
        const double* src = T + pl * 64 + (cg * 4) * 8 + row;
        v01 = make_double2(src[0], src[8]);
        v23 = make_double2(src[16], src[24]);
    }
    double* out = dst + kg_off(ld, row0 + row, col0 + cg * 4, pl);
    *reinterpret_cast<double2*>(out) = make_double2(sign * v01.x, sign * v01.y);
    *reinterpret_cast<double2*>(out + 2) = make_double2(sign * v23.x, sign * v23.y);
}

// ------------------------------------------------------------------------------------------------------------------------
// small fronts: one CTA per (front, system).  list[blockIdx.x] = front id.
template <int NW>
__global__ void __launch_bounds__(NW * 32)
mf_small_kernel(Tables tb, const int* __restrict__ list) {
    constexpr int NT = NW * 32;
    extern __shared__ __align__(16) unsigned char mf_smem[];
    const Front F = tb.fronts[list[blockIdx.x]];
    const int sys = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int fp = F.sp + F.up, nb = fp >> 3, npb = F.sp >> 3;
    const int nT = nb * (nb + 1) / 2;
    double* tiles = reinterpret_cast<double*>(mf_smem);
    double* raw = tiles + (size_t)nT * 128;
    double* mm = raw + 16 * (size_t)fp;
    double* nainv = mm + 16 * (size_t)fp;
    int* fail = reinterpret_cast<int*>(nainv + 128);
    constexpr int _pc = NW == 2 ? 0 : (NW == 4 ? 1 : (NW == 8 ? 2 : 3));      // profiler class
    long long _t0 = clock64();
    // loads that do not depend on anything else are issued first: their latency hides behind the zeroing and the extend-add
    const Chunk ch = tb.chunks[F.chunkPtr];
    const cplx* vals = tb.vals + (size_t)sys * tb.valStride;
    OrigEntry oe0{0, 0, -1};
    cplx ov0 = mk(0.0, 0.0);
    if (tid < F.nOrig) {
        oe0 = tb.orig[F.origPtr + tid];
        ov0 = oe0.src < 0 ? mk(1.0, 0.0) : vals[oe0.src];
    }
    for (int i = tid; i < nT * 64; i += NT) reinterpret_cast<double2*>(tiles)[i] = make_double2(0.0, 0.0);
    if (tid == 0) *fail = 0;
    __syncthreads();
    MF_PROF_MARK(0);
    auto addr = [&](int a, int b) {       // a >= b
        return tiles + (size_t)((a >> 3) * ((a >> 3) + 1) / 2 + (b >> 3)) * 128 + (a & 7) * 8 + (b & 7);
    };
    // extend-add of the children's update matrices, one child at a time (entries of one child never collide: plain adds, fixed
    // order -> deterministic).  The child's row map is staged in shared memory (the operand panel is still unused).
    const double* carena = tb.arena[(F.depth + 1) & 1] + (size_t)sys * tb.arenaStride[(F.depth + 1) & 1];
    int* relS = reinterpret_cast<int*>(raw);
    for (int c = 0; c < F.nChild; ++c) {
        const Front& C = tb.fronts[tb.children[F.childPtr + c]];
        const int cu = C.u, cup = C.up;
        const int ld = C.isBig ? C.sp + cup : cup, off = C.isBig ? C.sp : 0;
        const double* U = carena + C.frontOff;
        const int* rel = tb.rel + C.rowPtr;
        for (int i = tid; i < cu; i += NT) relS[i] = rel[i];
        __syncthreads();
        // work items = (column group of four, block of 32 rows) of the child's lower triangle, dealt round-robin; a warp issues
        // the loads of kBatch items (16-byte loads, 64 bytes per lane and item) before it scatters them, so that several
        // global-memory round trips are in flight.  (A 4-row x 8-column lane mapping avoids the shared-memory bank conflicts of
        // the scatter but measured slower: the phase is bound by the latency of these loads, not by the scatter.)
        const int ng = (cu + 3) >> 2, nrb = (cu + 31) >> 5;
        constexpr int kBatch = 4;
        for (int q0 = warp * kBatch; q0 < ng * nrb; q0 += NW * kBatch) {
            double2 r01[kBatch], r23[kBatch], i01[kBatch], i23[kBatch];
            int ii[kBatch], jgq[kBatch];
#pragma unroll
            for (int bq = 0; bq < kBatch; ++bq) {
                const int q = q0 + bq;
                const int jg = q / nrb, i = (q - jg * nrb) * 32 + lane;
                jgq[bq] = jg;
                ii[bq] = (q < ng * nrb && i < cu && i >= 4 * jg) ? i : -1;
                if (ii[bq] >= 0) {
                    const double* pr = U + kg_off(ld, off + i, off + 4 * jg, 0);
                    const double* pi = U + kg_off(ld, off + i, off + 4 * jg, 1);
                    r01[bq] = *reinterpret_cast<const double2*>(pr); r23[bq] = *reinterpret_cast<const double2*>(pr + 2);
                    i01[bq] = *reinterpret_cast<const double2*>(pi); i23[bq] = *reinterpret_cast<const double2*>(pi + 2);
                }
            }
#pragma unroll
            for (int bq = 0; bq < kBatch; ++bq) {
                const int i = ii[bq];
                if (i < 0) continue;
                const int jg = jgq[bq];
                const double re[4] = {r01[bq].x, r01[bq].y, r23[bq].x, r23[bq].y}, im[4] = {i01[bq].x, i01[bq].y, i23[bq].x, i23[bq].y};
                const int ri = relS[i];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    if (4 * jg + jj > i) break;
                    double* d = addr(ri, relS[4 * jg + jj]);
                    d[0] += re[jj];
                    d[64] += im[jj];
                }
            }
        }
        __syncthreads();
    }
    MF_PROF_MARK(2);
    // original matrix entries of the pivot columns
    if (tid < F.nOrig) {
        double* d = addr(oe0.lrow, oe0.lcol);
        d[0] += ov0.x;
        d[64] += ov0.y;
    }
    for (int e = tid + NT; e < F.nOrig; e += NT) {
        const OrigEntry oe = tb.orig[F.origPtr + e];
        const cplx v = oe.src < 0 ? mk(1.0, 0.0) : vals[oe.src];
        double* d = addr(oe.lrow, oe.lcol);
        d[0] += v.x;
        d[64] += v.y;
    }
    __syncthreads();
    MF_PROF_MARK(1);
    mf_sweep<NW>(tiles, raw, mm, nainv, fail, nb, npb, tb.prof ? tb.prof + 32 + _pc * 4 : nullptr);
    MF_PROF_MARK(4);
    if (tid == 0 && *fail) tb.status[sys] = kErrSingular;
    // outputs, tile by tile: G = -(pivot x pivot), M = update x pivot (factor arena), U = update x update (update arena, lower tiles)
    double* fac = tb.fac + (size_t)sys * tb.facStride;
    const int sp = F.sp, up = F.up, nub = nb - npb;
    auto tileP = [&](int I, int J) { return tiles + (size_t)(I * (I + 1) / 2 + J) * 128; };
    {
        double* G = fac + ch.gOff;
        for (int q = warp; q < npb * npb; q += NW) {
            const int I = q / npb, J = q - I * npb;
            if (I > J) mf_store_tile<0>(tileP(I, J), G, sp, I * 8, J * 8, -1.0, lane);
            else if (I == J) mf_store_tile<2>(tileP(I, I), G, sp, I * 8, I * 8, -1.0, lane);
            else mf_store_tile<1>(tileP(J, I), G, sp, I * 8, J * 8, -1.0, lane);
        }
    }
    if (up > 0) {
        double* M = fac + ch.mOff;
        for (int q = warp; q < nub * npb; q += NW) {
            const int I = q / npb, J = q - I * npb;
            mf_store_tile<0>(tileP(npb + I, J), M, up, I * 8, J * 8, 1.0, lane);
        }
        double* U = tb.arena[F.depth & 1] + (size_t)sys * tb.arenaStride[F.depth & 1] + F.frontOff;
        int I = 0, J = warp;
        while (J > I) { J -= I + 1; ++I; }
        for (int L = warp; L < nub * (nub + 1) / 2; L += NW) {
            mf_store_tile<0>(tileP(npb + I, npb + J), U, up, I * 8, J * 8, 1.0, lane);
            J += NW;
            while (J > I) { J -= I + 1; ++I; }
        }
    }
    MF_PROF_MARK(5);
    if (tb.prof && tid == 0) { atomicAdd(tb.prof + _pc * 8 + 6, 1ull); atomicAdd(tb.prof + _pc * 8 + 7, (unsigned long long)npb); }
}

// ------------------------------------------------------------------------------------------------------------------------
// large fronts, assembly.  (front, original entry) pairs of one depth: grid (ceil(n/256), nsys)
__global__ void mf_asm_orig_kernel(Tables tb, const int2* __restrict__ pairs, int n) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x, sys = blockIdx.y;
    if (idx >= n) return;
    const int2 pr = pairs[idx];
    const Front& F = tb.fronts[pr.x];
    const OrigEntry oe = tb.orig[pr.y];
    const cplx v = oe.src < 0 ? mk(1.0, 0.0) : tb.vals[(size_t)sys * tb.valStride + oe.src];
    double* A = tb.arena[F.depth & 1] + (size_t)sys * tb.arenaStride[F.depth & 1] + F.frontOff;
    const int fp = F.sp + F.up;
    A[kg_off(fp, oe.lrow, oe.lcol, 0)] += v.x;
    A[kg_off(fp, oe.lrow, oe.lcol, 1)] += v.y;
}
// extend-add of one child per parent (pass c handles the c-th child of every large front of the depth): pairs = (parent, child);
// grid (npairs, nsys, nseg)
__global__ void __launch_bounds__(256)
mf_asm_child_kernel(Tables tb, const int2* __restrict__ pairs) {
    const int2 pr = pairs[blockIdx.x];
    const int sys = blockIdx.y;
    const Front P = tb.fronts[pr.x], C = tb.fronts[pr.y];
    double* A = tb.arena[P.depth & 1] + (size_t)sys * tb.arenaStride[P.depth & 1] + P.frontOff;
    const double* U = tb.arena[C.depth & 1] + (size_t)sys * tb.arenaStride[C.depth & 1] + C.frontOff;
    const int fp = P.sp + P.up;
    const int ld = C.isBig ? C.sp + C.up : C.up, off = C.isBig ? C.sp : 0;
    const int* rel = tb.rel + C.rowPtr;
    const int ng = (C.u + 3) >> 2;
    for (int jg = blockIdx.z; jg < ng; jg += gridDim.z) {
        int rj[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) rj[jj] = (4 * jg + jj < C.u) ? rel[4 * jg + jj] : 0;
        for (int i = 4 * jg + threadIdx.x; i < C.u; i += 256) {
            const double* pr0 = U + kg_off(ld, off + i, off + 4 * jg, 0);
            const double* pi0 = U + kg_off(ld, off + i, off + 4 * jg, 1);
            const double2 r01 = *reinterpret_cast<const double2*>(pr0), r23 = *reinterpret_cast<const double2*>(pr0 + 2);
            const double2 i01 = *reinterpret_cast<const double2*>(pi0), i23 = *reinterpret_cast<const double2*>(pi0 + 2);
            const double re[4] = {r01.x, r01.y, r23.x, r23.y}, im[4] = {i01.x, i01.y, i23.x, i23.y};
            const int ri = rel[i];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                if (4 * jg + jj > i) break;
                A[kg_off(fp, ri, rj[jj], 0)] += re[jj];
                A[kg_off(fp, ri, rj[jj], 1)] += im[jj];
            }
        }
    }
}

// G of the diagonal block of chunk `c` of every listed large front: grid (nfronts, nsys), kInvWarps warps
constexpr int kInvWarps = 16;
__global__ void __launch_bounds__(kInvWarps * 32)
mf_inv_kernel(Tables tb, const int* __restrict__ list, int c) {
    constexpr int NT = kInvWarps * 32;
    extern __shared__ __align__(16) unsigned char mf_smem[];
    const Front F = tb.fronts[list[blockIdx.x]];
    const Chunk ch = tb.chunks[F.chunkPtr + c];
    const int sys = blockIdx.y, tid = threadIdx.x;
    const int sc = ch.p1 - ch.p0, nb = sc >> 3, fp = F.sp + F.up;
    const int nT = nb * (nb + 1) / 2;
    double* tiles = reinterpret_cast<double*>(mf_smem);
    double* raw = tiles + (size_t)nT * 128;
    double* mm = raw + 16 * (size_t)sc;
    double* nainv = mm + 16 * (size_t)sc;
    int* fail = reinterpret_cast<int*>(nainv + 128);
    const double* A = tb.arena[F.depth & 1] + (size_t)sys * tb.arenaStride[F.depth & 1] + F.frontOff;
    if (tid == 0) *fail = 0;
    // lower triangle of A[p0:p1][p0:p1] -> tiles (diagonal tiles mirrored)
    for (int idx = tid; idx < (sc >> 2) * sc; idx += NT) {
        const int jg = idx / sc, i = idx - jg * sc;
        if ((i >> 3) < (jg >> 1)) continue;                      // tile strictly above the diagonal
        const double* pr = A + kg_off(fp, ch.p0 + i, ch.p0 + 4 * jg, 0);
        const double* pi = A + kg_off(fp, ch.p0 + i, ch.p0 + 4 * jg, 1);
        const double2 r01 = *reinterpret_cast<const double2*>(pr), r23 = *reinterpret_cast<const double2*>(pr + 2);
        const double2 i01 = *reinterpret_cast<const double2*>(pi), i23 = *reinterpret_cast<const double2*>(pi + 2);
        const int I = i >> 3, J = jg >> 1;
        double* T = tiles + (size_t)(I * (I + 1) / 2 + J) * 128 + (i & 7) * 8 + (jg & 1) * 4;
        *reinterpret_cast<double2*>(T) = r01; *reinterpret_cast<double2*>(T + 2) = r23;
        *reinterpret_cast<double2*>(T + 64) = i01; *reinterpret_cast<double2*>(T + 66) = i23;
    }
    __syncthreads();
    for (int idx = tid; idx < nb * 64; idx += NT) {
        const int I = idx >> 6, r = (idx >> 3) & 7, cc = idx & 7;
        if (cc > r) {
            double* T = tiles + (size_t)(I * (I + 1) / 2 + I) * 128;
            T[r * 8 + cc] = T[cc * 8 + r];
            T[64 + r * 8 + cc] = T[64 + cc * 8 + r];
        }
    }
    __syncthreads();
    mf_sweep<kInvWarps>(tiles, raw, mm, nainv, fail, nb, nb);
    if (tid == 0 && *fail) tb.status[sys] = kErrSingular;
    double* G = tb.fac + (size_t)sys * tb.facStride + ch.gOff;
    for (int idx = tid; idx < (sc >> 2) * sc; idx += NT) {
        const int jg = idx / sc, i = idx - jg * sc;
        double re[4], im[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) { const cplx v = mf_tile_get(tiles, i, 4 * jg + jj); re[jj] = -v.x; im[jj] = -v.y; }
        double* pr = G + kg_off(sc, i, 4 * jg, 0);
        double* pi = G + kg_off(sc, i, 4 * jg, 1);
        *reinterpret_cast<double2*>(pr) = make_double2(re[0], re[1]); *reinterpret_cast<double2*>(pr + 2) = make_double2(re[2], re[3]);
        *reinterpret_cast<double2*>(pi) = make_double2(im[0], im[1]); *reinterpret_cast<double2*>(pi + 2) = make_double2(im[2], im[3]);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// batched complex GEMM  C (64x64 tile) = beta C + alpha A B^T  on the FP64 tensor cores; A (m x K), B (n x K), C (m x n) k-grouped.
// space: 0 / 1 = depth-parity arena, 2 = factor arena
struct GemmJob {
    int64_t aOff, bOff, cOff;
    int aSp, bSp, cSp;
    int ldA, rA, kA, ldB, rB, kB, ldC, rC, cC;
    int m, n, K;
    int lower, beta;
    double alpha;
};
struct GemmTile {
    int job, bi, bj;
};
constexpr int kGemmKC = 16, kGemmStages = 3, kGemmThreads = 256;
constexpr int kGemmStageDoubles = 2 * (kGemmKC / 4) * 2 * 64 * 4;      // A + B operand images of one stage
constexpr size_t kGemmSmemBytes = (size_t)kGemmStages * kGemmStageDoubles * sizeof(double);

__global__ void __launch_bounds__(kGemmThreads)
mf_gemm_kernel(Tables tb, const GemmJob* __restrict__ jobs, const GemmTile* __restrict__ tilesList) {
    extern __shared__ __align__(16) unsigned char mf_smem[];
    double* sm = reinterpret_cast<double*>(mf_smem);
    const GemmTile gt = tilesList[blockIdx.x];
    const GemmJob jb = jobs[gt.job];
    const int sys = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = lane >> 2, t = lane & 3;
    const int wr = warp >> 1, wc = warp & 1;          // warp tile: rows [16 wr, +16), cols [32 wc, +32)
    auto base = [&](int sp, int64_t off) -> double* {
        return (sp == 2 ? tb.fac + (size_t)sys * tb.facStride : tb.arena[sp] + (size_t)sys * tb.arenaStride[sp]) + off;
    };
    const double* A = base(jb.aSp, jb.aOff);
    const double* B = base(jb.bSp, jb.bOff);
    double* C = base(jb.cSp, jb.cOff);
    const int r0 = gt.bi * 64, c0 = gt.bj * 64;
    const int nk = (jb.K + kGemmKC - 1) / kGemmKC;
    // stage image: [A|B][group][plane][64][4]
    auto load_stage = [&](int ks, int st) {
        double* dst = sm + (size_t)st * kGemmStageDoubles;
        const int k0 = ks * kGemmKC;
#pragma unroll
        for (int it = 0; it < 2 * (kGemmKC / 4) * 2 * 128 / kGemmThreads; ++it) {
            const int c = tid + it * kGemmThreads;
            const int which = c / ((kGemmKC / 4) * 2 * 128);         // 0: A, 1: B
            const int cc = c - which * ((kGemmKC / 4) * 2 * 128);
            const int grp = cc >> 8, pl = (cc >> 7) & 1, row = (cc & 127) >> 1, half = cc & 1;
            const int k = k0 + 4 * grp;
            const double* src;
            bool ok;
            if (which == 0) {
                ok = (r0 + row < jb.m) && (k < jb.K);
                src = A + kg_off(jb.ldA, jb.rA + r0 + row, jb.kA + k, pl) + half * 2;
            } else {
                ok = (c0 + row < jb.n) && (k < jb.K);
                src = B + kg_off(jb.ldB, jb.rB + c0 + row, jb.kB + k, pl) + half * 2;
            }
            cp_async16(dst + (size_t)c * 2, ok ? src : A, ok ? 16u : 0u);
        }
        cp_async_commit();
    };
    double cre[2][4][2], cim[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) cre[a][b][0] = cre[a][b][1] = cim[a][b][0] = cim[a][b][1] = 0.0;
    for (int s = 0; s < kGemmStages - 1; ++s) {
        if (s < nk) load_stage(s, s); else cp_async_commit();
    }
    for (int ks = 0; ks < nk; ++ks) {
        cp_async_wait<kGemmStages - 2>();
        __syncthreads();
        if (ks + kGemmStages - 1 < nk) load_stage(ks + kGemmStages - 1, (ks + kGemmStages - 1) % kGemmStages); else cp_async_commit();
        const double* sA = sm + (size_t)(ks % kGemmStages) * kGemmStageDoubles;
        const double* sB = sA + (kGemmKC / 4) * 2 * 64 * 4;
#pragma unroll
        for (int grp = 0; grp < kGemmKC / 4; ++grp) {
            double are[2], aim[2], nai[2], bre[4], bim[4];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                are[a] = sA[((grp * 2 + 0) * 64 + wr * 16 + a * 8 + g) * 4 + t];
                aim[a] = sA[((grp * 2 + 1) * 64 + wr * 16 + a * 8 + g) * 4 + t];
                nai[a] = -aim[a];
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                bre[b] = sB[((grp * 2 + 0) * 64 + wc * 32 + b * 8 + g) * 4 + t];
                bim[b] = sB[((grp * 2 + 1) * 64 + wc * 32 + b * 8 + g) * 4 + t];
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    dmma884(cre[a][b], are[a], bre[b]);
                    dmma884(cim[a][b], are[a], bim[b]);
                }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    dmma884(cre[a][b], nai[a], bim[b]);
                    dmma884(cim[a][b], aim[a], bre[b]);
                }
        }
    }
    cp_async_wait<0>();
    // epilogue: C fragments straight to / from global memory (16-byte pairs)
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int rb = r0 + wr * 16 + a * 8, cb = c0 + wc * 32 + b * 8;
            if (rb >= jb.m || cb >= jb.n) continue;
            if (jb.lower && jb.cC + cb > jb.rC + rb) continue;
            double* pr = C + kg_off(jb.ldC, jb.rC + rb + g, jb.cC + cb + 2 * t, 0);
            double* pi = C + kg_off(jb.ldC, jb.rC + rb + g, jb.cC + cb + 2 * t, 1);
            double2 vr = make_double2(jb.alpha * cre[a][b][0], jb.alpha * cre[a][b][1]);
            double2 vi = make_double2(jb.alpha * cim[a][b][0], jb.alpha * cim[a][b][1]);
            if (jb.beta) {
                const double2 orr = *reinterpret_cast<const double2*>(pr), oi = *reinterpret_cast<const double2*>(pi);
                vr.x += orr.x; vr.y += orr.y; vi.x += oi.x; vi.y += oi.y;
            }
            *reinterpret_cast<double2*>(pr) = vr;
            *reinterpret_cast<double2*>(pi) = vi;
        }
}

// ------------------------------------------------------------------------------------------------------------------------
// values of the MT stencil systems: vals[sys] = [dr + i omega dm | e1 | e2]  (pattern: mf_grid_entries).  grid (ceil(3N/256), nsys)
__global__ void mf_mt_vals_kernel(int N, const MtValSys* __restrict__ sysv, cplx* __restrict__ vals, int64_t valStride) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x, sys = blockIdx.y;
    if (e >= 3 * N) return;
    const MtValSys sv = sysv[sys];
    cplx v;
    if (e < N) v = mk(sv.planes[e], sv.omega * sv.planes[N + e]);
    else v = mk(sv.planes[(size_t)N + e], 0.0);      // e in [N,2N): e1 = planes[2N + q] ; [2N,3N): e2 = planes[3N + q]
    vals[(size_t)sys * valStride + e] = v;
}

// ------------------------------------------------------------------------------------------------------------------------
// solves.  One CTA per (front, right-hand side); vec = blockIdx.y indexes (system, rhs): sys = vec / nrhs.
//   v   : [nvec][Np]  forward: pivot parts of the eliminated rhs; backward: overwritten by the solution (padded permuted numbering)
//   upd : [nvec][updEntries] update vectors handed from the children to their parent
struct SolveArgs {
    const cplx* B;      // right-hand sides, original numbering, vector `vec` at B + vec*ldb
    cplx* X;            // solutions, same layout (may alias B)
    int64_t ldb, ldx;
    cplx* v;
    cplx* upd;
    int64_t Np, updEntries;
    int nrhs;
};
constexpr int kSolveMfThreads = 256;

__global__ void __launch_bounds__(kSolveMfThreads)
mf_fwd_kernel(Tables tb, SolveArgs sa, const int* __restrict__ list) {
    extern __shared__ __align__(16) unsigned char mf_smem[];
    cplx* w = reinterpret_cast<cplx*>(mf_smem);
    const Front F = tb.fronts[list[blockIdx.x]];
    const int vec = blockIdx.y, sys = vec / sa.nrhs, tid = threadIdx.x;
    const int fp = F.sp + F.up;
    const cplx* b = sa.B + (size_t)vec * sa.ldb;
    cplx* v = sa.v + (size_t)vec * sa.Np;
    cplx* upd = sa.upd + (size_t)vec * sa.updEntries;
    for (int i = tid; i < fp; i += kSolveMfThreads) {
        cplx x = mk(0.0, 0.0);
        if (i < F.s) x = b[tb.pos2orig[F.cbp + i]];
        w[i] = x;
    }
    __syncthreads();
    for (int c = 0; c < F.nChild; ++c) {
        const Front& C = tb.fronts[tb.children[F.childPtr + c]];
        const int* rel = tb.rel + C.rowPtr;
        const cplx* uv = upd + C.updOff;
        const int cu = C.u;
        for (int i = tid; i < cu; i += kSolveMfThreads) w[rel[i]] += uv[i];
        __syncthreads();
    }
    const double* fac = tb.fac + (size_t)sys * tb.facStride;
    for (int c = 0; c < F.nChunk; ++c) {
        const Chunk ch = tb.chunks[F.chunkPtr + c];
        const int sc = ch.p1 - ch.p0, mr = fp - ch.p1;
        const double* M = fac + ch.mOff;
        for (int i = tid; i < mr; i += kSolveMfThreads) {
            cplx acc = mk(0.0, 0.0);
            for (int kg = 0; kg < (sc >> 2); ++kg) {
                const double* pr = M + kg_off(mr, i, 4 * kg, 0);
                const double* pi = M + kg_off(mr, i, 4 * kg, 1);
                const double2 r01 = *reinterpret_cast<const double2*>(pr), r23 = *reinterpret_cast<const double2*>(pr + 2);
                const double2 i01 = *reinterpret_cast<const double2*>(pi), i23 = *reinterpret_cast<const double2*>(pi + 2);
                const cplx* wk = w + ch.p0 + 4 * kg;
                cfma(acc, mk(r01.x, i01.x), wk[0]);
                cfma(acc, mk(r01.y, i01.y), wk[1]);
                cfma(acc, mk(r23.x, i23.x), wk[2]);
                cfma(acc, mk(r23.y, i23.y), wk[3]);
            }
            w[ch.p1 + i] -= acc;
        }
        __syncthreads();
    }
    for (int i = tid; i < F.sp; i += kSolveMfThreads) v[F.cbp + i] = w[i];
    for (int i = tid; i < F.up; i += kSolveMfThreads) upd[F.updOff + i] = w[F.sp + i];
}

__global__ void __launch_bounds__(kSolveMfThreads)
mf_bwd_kernel(Tables tb, SolveArgs sa, const int* __restrict__ list) {
    extern __shared__ __align__(16) unsigned char mf_smem[];
    const Front F = tb.fronts[list[blockIdx.x]];
    const int vec = blockIdx.y, sys = vec / sa.nrhs, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fp = F.sp + F.up;
    cplx* xf = reinterpret_cast<cplx*>(mf_smem);        // [fp]
    cplx* tmp = xf + fp;                                 // [kChunkMax or sp]
    cplx* v = sa.v + (size_t)vec * sa.Np;
    cplx* x = sa.X + (size_t)vec * sa.ldx;
    const int* rows = tb.rows + F.rowPtr;
    for (int i = tid; i < fp; i += kSolveMfThreads) {
        cplx val = mk(0.0, 0.0);
        if (i < F.sp) val = v[F.cbp + i];
        else if (i - F.sp < F.u) val = v[rows[i - F.sp]];
        xf[i] = val;
    }
    __syncthreads();
    const double* fac = tb.fac + (size_t)sys * tb.facStride;
    constexpr int NWS = kSolveMfThreads / 32;
    for (int c = F.nChunk - 1; c >= 0; --c) {
        const Chunk ch = tb.chunks[F.chunkPtr + c];
        const int sc = ch.p1 - ch.p0, mr = fp - ch.p1;
        const double* G = fac + ch.gOff;
        const double* M = fac + ch.mOff;
        // x1[k] = sum_j G[j][k] w1[j] - sum_i M[i][k] x2[i] : one warp per group of four k
        for (int kg = warp; kg < (sc >> 2); kg += NWS) {
            cplx a0 = mk(0.0, 0.0), a1 = a0, a2 = a0, a3 = a0;
            for (int j = lane; j < sc; j += 32) {
                const double* pr = G + kg_off(sc, j, 4 * kg, 0);
                const double* pi = G + kg_off(sc, j, 4 * kg, 1);
                const double2 r01 = *reinterpret_cast<const double2*>(pr), r23 = *reinterpret_cast<const double2*>(pr + 2);
                const double2 i01 = *reinterpret_cast<const double2*>(pi), i23 = *reinterpret_cast<const double2*>(pi + 2);
                const cplx wj = xf[ch.p0 + j];
                cfma(a0, mk(r01.x, i01.x), wj); cfma(a1, mk(r01.y, i01.y), wj);
                cfma(a2, mk(r23.x, i23.x), wj); cfma(a3, mk(r23.y, i23.y), wj);
            }
            for (int i = lane; i < mr; i += 32) {
                const double* pr = M + kg_off(mr, i, 4 * kg, 0);
                const double* pi = M + kg_off(mr, i, 4 * kg, 1);
                const double2 r01 = *reinterpret_cast<const double2*>(pr), r23 = *reinterpret_cast<const double2*>(pr + 2);
                const double2 i01 = *reinterpret_cast<const double2*>(pi), i23 = *reinterpret_cast<const double2*>(pi + 2);
                const cplx xi = -xf[ch.p1 + i];
                cfma(a0, mk(r01.x, i01.x), xi); cfma(a1, mk(r01.y, i01.y), xi);
                cfma(a2, mk(r23.x, i23.x), xi); cfma(a3, mk(r23.y, i23.y), xi);
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                a0.x += __shfl_xor_sync(0xffffffffu, a0.x, off); a0.y += __shfl_xor_sync(0xffffffffu, a0.y, off);
                a1.x += __shfl_xor_sync(0xffffffffu, a1.x, off); a1.y += __shfl_xor_sync(0xffffffffu, a1.y, off);
                a2.x += __shfl_xor_sync(0xffffffffu, a2.x, off); a2.y += __shfl_xor_sync(0xffffffffu, a2.y, off);
                a3.x += __shfl_xor_sync(0xffffffffu, a3.x, off); a3.y += __shfl_xor_sync(0xffffffffu, a3.y, off);
            }
            if (lane == 0) { tmp[4 * kg] = a0; tmp[4 * kg + 1] = a1; tmp[4 * kg + 2] = a2; tmp[4 * kg + 3] = a3; }
        }
        __syncthreads();
        for (int k = tid; k < sc; k += kSolveMfThreads) xf[ch.p0 + k] = tmp[k];
        __syncthreads();
    }
    for (int i = tid; i < F.sp; i += kSolveMfThreads) {
        v[F.cbp + i] = xf[i];
        if (i < F.s) x[tb.pos2orig[F.cbp + i]] = xf[i];
    }
}


// Small fronts (single chunk, fp <= kSolveSmallMax): one WARP per (front, right-hand side), eight fronts per CTA, warp-level
// synchronisation only.  list[blockIdx.x * 8 + warp] = front id.
constexpr int kSolveSmallMax = 64;
constexpr int kSolveWarpsPerCta = 8;
__global__ void __launch_bounds__(kSolveWarpsPerCta * 32)
mf_fwd_warp_kernel(Tables tb, SolveArgs sa, const int* __restrict__ list, int n) {
    __shared__ cplx wsh[kSolveWarpsPerCta][kSolveSmallMax];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fi = blockIdx.x * kSolveWarpsPerCta + warp;
    if (fi >= n) return;
    const Front& F = tb.fronts[list[fi]];
    const int vec = blockIdx.y, sys = vec / sa.nrhs;
    const int sp = F.sp, up = F.up, fp = sp + up, fs = F.s, cbp = F.cbp;
    cplx* w = wsh[warp];
    const cplx* b = sa.B + (size_t)vec * sa.ldb;
    cplx* v = sa.v + (size_t)vec * sa.Np;
    cplx* upd = sa.upd + (size_t)vec * sa.updEntries;
    for (int i = lane; i < fp; i += 32) w[i] = i < fs ? b[tb.pos2orig[cbp + i]] : mk(0.0, 0.0);
    __syncwarp();
    const int nChild = F.nChild, childPtr = F.childPtr;
    for (int c = 0; c < nChild; ++c) {
        const Front& C = tb.fronts[tb.children[childPtr + c]];
        const int* rel = tb.rel + C.rowPtr;
        const cplx* uv = upd + C.updOff;
        const int cu = C.u;
        for (int i = lane; i < cu; i += 32) w[rel[i]] += uv[i];
        __syncwarp();
    }
    if (up > 0) {
        const double* M = tb.fac + (size_t)sys * tb.facStride + tb.chunks[F.chunkPtr].mOff;
        for (int i = lane; i < up; i += 32) {
            cplx acc = mk(0.0, 0.0);
            for (int kg = 0; kg < (sp >> 2); ++kg) {
                const double* pr = M + kg_off(up, i, 4 * kg, 0);
                const double* pi = M + kg_off(up, i, 4 * kg, 1);
                const double2 r01 = *reinterpret_cast<const double2*>(pr), r23 = *reinterpret_cast<const double2*>(pr + 2);
                const double2 i01 = *reinterpret_cast<const double2*>(pi), i23 = *reinterpret_cast<const double2*>(pi + 2);
                const cplx* wk = w + 4 * kg;
                cfma(acc, mk(r01.x, i01.x), wk[0]);
                cfma(acc, mk(r01.y, i01.y), wk[1]);
                cfma(acc, mk(r23.x, i23.x), wk[2]);
                cfma(acc, mk(r23.y, i23.y), wk[3]);
            }
            upd[F.updOff + i] = w[sp + i] - acc;
        }
    }
    for (int i = lane; i < sp; i += 32) v[cbp + i] = w[i];
}

__global__ void __launch_bounds__(kSolveWarpsPerCta * 32)
mf_bwd_warp_kernel(Tables tb, SolveArgs sa, const int* __restrict__ list, int n) {
    __shared__ cplx xsh[kSolveWarpsPerCta][kSolveSmallMax];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fi = blockIdx.x * kSolveWarpsPerCta + warp;
    if (fi >= n) return;
    const Front& F = tb.fronts[list[fi]];
    const int vec = blockIdx.y, sys = vec / sa.nrhs;
    const int sp = F.sp, up = F.up, fp = sp + up, fs = F.s, fu = F.u, cbp = F.cbp;
    cplx* xf = xsh[warp];
    cplx* v = sa.v + (size_t)vec * sa.Np;
    cplx* x = sa.X + (size_t)vec * sa.ldx;
    const int* rows = tb.rows + F.rowPtr;
    for (int i = lane; i < fp; i += 32) {
        cplx val = mk(0.0, 0.0);
        if (i < sp) val = v[cbp + i];
        else if (i - sp < fu) val = v[rows[i - sp]];
        xf[i] = val;
    }
    __syncwarp();
    const Chunk ch = tb.chunks[F.chunkPtr];
    const double* G = tb.fac + (size_t)sys * tb.facStride + ch.gOff;
    const double* M = tb.fac + (size_t)sys * tb.facStride + ch.mOff;
    // x1[k] = sum_j G[j][k] w1[j] - sum_i M[i][k] x2[i] : one lane per k (four lanes share a 32-byte sector); when the front has
    // fewer than 32 pivots the rows are split over 32 / width sub-groups of lanes and combined with shuffles
    const int width = sp <= 8 ? 8 : (sp <= 16 ? 16 : 32), nparts = 32 / width, part = lane / width, kl = lane - part * width;
    for (int k0 = 0; k0 < sp; k0 += width) {
        const int k = k0 + kl;
        cplx acc = mk(0.0, 0.0);
        if (k < sp) {
            const double* gr = G + kg_off(sp, 0, k, 0);
            const double* gi = G + kg_off(sp, 0, k, 1);
            for (int j = part; j < sp; j += nparts) cfma(acc, mk(gr[4 * j], gi[4 * j]), xf[j]);
            const double* mr = M + kg_off(up, 0, k, 0);
            const double* mi = M + kg_off(up, 0, k, 1);
            for (int i = part; i < up; i += nparts) cfma(acc, mk(-mr[4 * i], -mi[4 * i]), xf[sp + i]);
        }
        for (int off = width; off < 32; off <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
        if (k < sp && part == 0) {
            v[cbp + k] = acc;
            if (k < fs) x[tb.pos2orig[cbp + k]] = acc;
        }
    }
}

}  // namespace mf
}  // namespace hmcmt
