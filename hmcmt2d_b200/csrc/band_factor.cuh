// Batched complex-symmetric banded block-LDL^T factorisation for sm_100a.
//
// Replaces the per-frequency sparse direct factorisation the reference obtains from
// `factorMUMPS(Aii,1)` / `lu(Aii)` (mt2DTE.jl:47-55, mt2DTM.jl:46-54, MUMPSfuncs.jl:24-39).
//
// Algorithm (validated in tools/proto/tile_band_proto.py): the matrix is kept in the internal
// "fast-axis" ordering with half-bandwidth b.  A sliding window of T x T tiles of 8x8 complex
// entries (8T >= b+8) lives entirely in REGISTERS as FP64 tensor-core accumulator fragments:
// tile slots are addressed circularly (global tile-block beta -> slot beta mod T) and the
// unordered slot pair {I,J} always holds exactly one live lower-triangle tile, so T(T+1)/2
// register tiles are 100 % occupied and nothing ever moves.  One macro-step eliminates an
// 8-column panel:
//     S  <-  S - raw * A11^{-1} * raw^T            (pivot-free block elimination)
//   * the tile warps form M' = raw * (-A11^{-1}) and apply the rank-8 update with
//     mma.sync.m8n8k4.f64 (DMMA.8x8x4, measured full-rate 37.1 TF/s on B200); the next pivot
//     column is updated first and published to shared memory;
//   * the "factor warp" runs one panel AHEAD of them: it applies the current panel's update to the
//     next 8x8 pivot block itself (16 DMMAs), inverts it in registers (Gauss-Jordan through warp
//     shuffles, pivot-free), forms z = A11^{-1} y_p for the fused right-hand side and streams the
//     finished panel to HBM with TMA bulk stores — the sequential pivot chain never stalls the tiles;
//   * recycled slots are refilled from the 5-point stencil planes (assembly is fused: the matrix is
//     never written to memory); those rows are prefetched into a shared-memory ring two steps ahead.
// The stored factor is { raw_s (8T x 8), [A11_s^{-1} (8x8) | z_s (8)] } per macro-step; it serves the
// forward solve (back-substitution fused below) and the adjoint solve (band_solve.cuh).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace hmcmt {

constexpr int TS = 8;         // tile size == DMMA m,n
constexpr int AZ = 72;        // complex entries per macro-step in the ainv stream: A11^{-1} (64) + z (8)

// Per-system description for a batched launch (one CTA per system).
struct BandSys {
    // stencil provider (internal ordering, see mt_kernels.cuh): A[g][g] = dr[g] + i*omega*dm[g],
    // A[g][g-1] = e1[g] (0 at line starts), A[g][g-nf] = e2[g] (0 on the first line)
    const double* dr;
    const double* dm;
    const double* e1;
    const double* e2;
    double omega;
    // dense lower-band provider (generic matrices through the MUMPS-shim ABI): band[g*(b+1)+d] = A[g][g-d]
    const cplx* band;
    const cplx* rhs;     // fused forward right-hand side (internal ordering, length N) or nullptr
    double* panels;      // [S][16*R] doubles: re/im x kk x R x 4 (exactly the smem operand layout)
    cplx* ainvz;         // [S][72]: A11^{-1} row-major, then z = A11^{-1} * (forward-eliminated rhs block)
    cplx* x;             // [N] solution of the fused system — only if rhs != nullptr
    int* status;         // 0 ok, -10 zero/NaN pivot block
};

__host__ __device__ constexpr int band_T_for(int b) { return (b + 7) / 8 + 1; }
__host__ __device__ constexpr int panel_doubles(int T) { return 16 * TS * T; }

template <int T>
struct FactorCfg {
    static constexpr int NT = T * (T + 1) / 2;
    // tile {X,Y} belongs to warp (X+Y) mod (T+1): every warp owns exactly T/2 tiles and at most ONE tile of any
    // pivot column, so recycle / look-ahead / trailing-update roles are balanced and found by one table lookup
    static_assert(T % 2 == 0, "even tile windows only");
    static constexpr int NW = T + 1;
    static constexpr int TPW = T / 2;
    static constexpr int R = TS * T;
    static constexpr int NTHREADS = (NW + 1) * 32;
};

constexpr int kRing = 4;       // prefetch ring depth (recycle rows / rhs), filled kPre steps ahead
constexpr int kPre = 2;

template <int T>
struct FactorSmem {
    static constexpr int R = TS * T;
    static constexpr int kBackStages = 5;
    // [buf][re/im][kk][r][t]  — raw[buf] is one contiguous panel image (16*R doubles)
    double raw[2][2][2][R][4];
    double m[2][2][2][R][4];
    double nainv[2][2][2][8][4];   // -A11^{-1} in B-fragment layout [buf][re/im][kk][n][t]
    double mscr[2][2][8][4];       // factor warp: M' of the next pivot block (A-fragment layout)
    double dnext[2][2][8][8];      // [buf][re/im][row][col]: next-next pivot block, updated through the current panel
    cplx ainvz[2][AZ];             // plain A11^{-1} (row-major) + z, staged for the TMA store
    cplx y[R];                     // circular window of the forward-eliminated rhs / back-substituted x
    double ringRow[kRing][4][8];   // stencil planes (dr, dm, e1, e2) of the rows entering the window
    cplx ringRhs[kRing][8];
    cplx part[16][8];              // back-substitution partial sums
    double stage[kBackStages][2][2][R][4];   // TMA landing buffers of the fused back-substitution
    cplx stageAZ[kBackStages][AZ];
    unsigned char tI[FactorCfg<T>::NT], tJ[FactorCfg<T>::NT];    // [warp*TPW + i]
    signed char li[FactorCfg<T>::NW][T];                          // local index of the tile of warp w touching block b (-1: none)
    uint64_t mbar[kBackStages];
    int fail;
};

enum { BAR_RAW = 1, BAR_INV = 2, BAR_M = 3 };

struct EntryProvider {
    const double *dr, *dm, *e1, *e2;
    const cplx* band;
    double omega;
    int N, nf, b;
    // direct (global-memory) evaluation, hi >= lo
    __device__ __forceinline__ cplx get(int hi, int lo) const {
        int d = hi - lo;
        if (hi >= N) return mk(d == 0 ? 1.0 : 0.0, 0.0);      // identity padding past the end
        if (band) return d <= b ? band[(size_t)hi * (b + 1) + d] : mk(0.0, 0.0);
        if (d == 0) return mk(dr[hi], omega * dm[hi]);
        if (d == 1) return mk(e1[hi], 0.0);
        if (d == nf) return mk(e2[hi], 0.0);
        return mk(0.0, 0.0);
    }
    // rows of one 8-row block staged in shared memory: row[plane][hi & 7]
    __device__ __forceinline__ cplx get_staged(const double (*row)[8], int hi, int lo) const {
        int d = hi - lo, r = hi & 7;
        if (d == 0) return mk(row[0][r], omega * row[1][r]);
        if (d == 1) return mk(row[2][r], 0.0);
        if (d == nf) return mk(row[3][r], 0.0);
        return mk(0.0, 0.0);
    }
};

// One CTA per system.  grid = nsys, block = FactorCfg<T>::NTHREADS, dynamic smem = sizeof(FactorSmem<T>)
template <int T>
__global__ void __launch_bounds__(FactorCfg<T>::NTHREADS, 1)
band_factor_kernel(const BandSys* __restrict__ systems, int N, int nf, int b) {
    using Cfg = FactorCfg<T>;
    constexpr int NT = Cfg::NT, NW = Cfg::NW, TPW = Cfg::TPW, R = Cfg::R, NTHR = Cfg::NTHREADS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FactorSmem<T>& sm = *reinterpret_cast<FactorSmem<T>*>(smem_raw);

    const BandSys sys = systems[blockIdx.x];
    const int S = (N + TS - 1) / TS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    EntryProvider prov{sys.dr, sys.dm, sys.e1, sys.e2, sys.band, sys.omega, N, nf, b};
    const bool staged = (sys.band == nullptr);      // stencil rows come through the smem ring

    // tile tables: warp w owns the unordered slot pairs {X,Y}, X <= Y, with (X+Y) mod NW == w
    for (int w = tid; w < NW; w += NTHR) {
        int cnt = 0;
        for (int bb = 0; bb < T; ++bb) sm.li[w][bb] = -1;
        for (int X = 0; X < T; ++X) {
            int Y = w - X; if (Y < 0) Y += NW;
            if (Y >= T || Y < X) continue;
            sm.tI[w * TPW + cnt] = (unsigned char)X;
            sm.tJ[w * TPW + cnt] = (unsigned char)Y;
            sm.li[w][X] = (signed char)cnt;
            sm.li[w][Y] = (signed char)cnt;
            ++cnt;
        }
    }
    for (int i = tid; i < R; i += NTHR) {
        cplx v = mk(0.0, 0.0);
        if (sys.rhs && i < N) v = sys.rhs[i];
        sm.y[i] = v;
    }
    // prefetch ring: entry q holds the rows of global block (T + q) for q < kPre (consumed at steps 0..kPre-1)
    auto ring_fetch = [&](int beta, int l, double& rowv, cplx& rhsv) {     // one warp, lane l
        int gr = beta * TS + (l & 7), pl = l >> 3;
        rowv = 0.0;
        if (gr < N) {
            if (staged) rowv = (pl == 0) ? sys.dr[gr] : (pl == 1) ? sys.dm[gr] : (pl == 2) ? sys.e1[gr] : sys.e2[gr];
        } else if (pl == 0) rowv = 1.0;
        rhsv = mk(0.0, 0.0);
        if (sys.rhs && l < 8 && gr < N) rhsv = sys.rhs[gr];
    };
    if (warp == 0) {
        for (int q = 0; q < kPre; ++q) {
            double rv; cplx hv;
            ring_fetch(T + q, lane, rv, hv);
            sm.ringRow[q % kRing][lane >> 3][lane & 7] = rv;
            if (lane < 8) sm.ringRhs[q % kRing][lane] = hv;
        }
    }
    if (tid == 0) sm.fail = 0;
    __syncthreads();

    if (warp < NW) {
        // =============================== tile warps ===============================
        double cre[TPW][2], cim[TPW][2];
        int tI[TPW], tJ[TPW];
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
            const bool valid = true;
            tI[i] = sm.tI[warp * TPW + i];
            tJ[i] = sm.tJ[warp * TPW + i];
            // initial window: slot block X holds global block X
            cre[i][0] = cre[i][1] = cim[i][0] = cim[i][1] = 0.0;
            if (valid) {
                int gi = tI[i] * TS + g;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int gj = tJ[i] * TS + 2 * t + e;
                    cplx v = prov.get(max(gi, gj), min(gi, gj));
                    cre[i][e] = v.x;
                    cim[i][e] = v.y;
                }
            }
        }
        // dump of the tiles touching slot block `q` into raw[nb] (panel columns = block q)
        auto dump = [&](int i, int q, int nb) {
            if (tJ[i] == q) {   // rows <-> I, cols <-> pivot  (also the diagonal tile)
                int r = tI[i] * TS + g;
                *reinterpret_cast<double2*>(&sm.raw[nb][0][t >> 1][r][(t & 1) * 2]) = make_double2(cre[i][0], cre[i][1]);
                *reinterpret_cast<double2*>(&sm.raw[nb][1][t >> 1][r][(t & 1) * 2]) = make_double2(cim[i][0], cim[i][1]);
            } else {            // tI == q: rows <-> pivot, cols <-> J : transposed
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int r = tJ[i] * TS + 2 * t + e;
                    sm.raw[nb][0][g >> 2][r][g & 3] = cre[i][e];
                    sm.raw[nb][1][g >> 2][r][g & 3] = cim[i][e];
                }
            }
        };
        auto dump_diag = [&](int i, int nb) {      // plain 8x8 image of a diagonal tile for the factor warp
            *reinterpret_cast<double2*>(&sm.dnext[nb][0][g][2 * t]) = make_double2(cre[i][0], cre[i][1]);
            *reinterpret_cast<double2*>(&sm.dnext[nb][1][g][2 * t]) = make_double2(cim[i][0], cim[i][1]);
        };
        // DMMA latency on B200 is ~138 cycles (tools/ubench): dependent accumulation chains are kept short
        // (4 independent 2-link chains per tile here) and, in pass 2, interleaved across tiles.
        auto update = [&](int i, int buf) {
            const int ra = tI[i] * TS + g, rb = tJ[i] * TS + g;
            double t1[2] = {0.0, 0.0}, t2[2] = {0.0, 0.0};
            double are[2], aim[2], bre[2], bim[2];
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                are[kk] = sm.m[buf][0][kk][ra][t]; aim[kk] = sm.m[buf][1][kk][ra][t];
                bre[kk] = sm.raw[buf][0][kk][rb][t]; bim[kk] = sm.raw[buf][1][kk][rb][t];
            }
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                dmma884(cre[i], are[kk], bre[kk]);
                dmma884(cim[i], are[kk], bim[kk]);
                dmma884(t1, -aim[kk], bim[kk]);
                dmma884(t2, aim[kk], bre[kk]);
            }
            cre[i][0] += t1[0]; cre[i][1] += t1[1];
            cim[i][0] += t2[0]; cim[i][1] += t2[1];
        };
        // prologue: publish panel 0 (tiles touching block 0) and the pivot block of panel 1
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
            if (tI[i] == 0) dump(i, 0, 0);      // tI==0 covers every tile touching block 0 (I<=J)
            if (T > 1 && tI[i] == 1 && tJ[i] == 1) dump_diag(i, 1);
        }
        bar_arrive(BAR_RAW, NTHR);

        // fused forward elimination of the rhs for block X: y_X -= raw_X z   (z = A11^{-1} y_p from the factor warp)
        auto y_update = [&](int X, int buf) {
            if (lane < 8) {
                const int ry = X * TS + lane;
                cplx acc = sm.y[ry];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    cplx rv = mk(sm.raw[buf][0][k >> 2][ry][k & 3], sm.raw[buf][1][k >> 2][ry][k & 3]);
                    cfma(acc, -rv, sm.ainvz[buf][64 + k]);
                }
                sm.y[ry] = acc;
            }
        };
        for (int s = 0; s < S; ++s) {
            const int p = s % T, p1 = (s + 1) % T, p2 = (s + 2) % T, buf = s & 1;
            // roles of this warp's tiles in this step (at most one tile per pivot column)
            const int ip = sm.li[warp][p], ip1 = sm.li[warp][p1];
            int idg = -1;                                   // the pivot block of panel s+2, if this warp owns it
            if (s + 1 < S) { int c = sm.li[warp][p2]; if (c >= 0 && 2 * p2 == ((2 * p2 >= NW) ? warp + NW : warp)) idg = c; }
            bar_sync(BAR_INV, NTHR);                       // raw(s) complete, -A11^{-1}(s) and z(s) published
            if (sm.fail) break;
            // M'_X = raw_X * (-A11^{-1}) for slot block X != p (one block per warp)
            int myX = -1;
            if (warp < T - 1) {
                int X = p + 1 + warp; if (X >= T) X -= T;
                myX = X;
                double mre[2] = {0.0, 0.0}, mim[2] = {0.0, 0.0}, mr2[2] = {0.0, 0.0}, mi2[2] = {0.0, 0.0};
                const int r = X * TS + g;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    double are = sm.raw[buf][0][kk][r][t], aim = sm.raw[buf][1][kk][r][t];
                    double bre = sm.nainv[buf][0][kk][g][t], bim = sm.nainv[buf][1][kk][g][t];
                    dmma884(mre, are, bre);
                    dmma884(mim, are, bim);
                    dmma884(mr2, -aim, bim);
                    dmma884(mi2, aim, bre);
                }
                mre[0] += mr2[0]; mre[1] += mr2[1]; mim[0] += mi2[0]; mim[1] += mi2[1];
                *reinterpret_cast<double2*>(&sm.m[buf][0][t >> 1][r][(t & 1) * 2]) = make_double2(mre[0], mre[1]);
                *reinterpret_cast<double2*>(&sm.m[buf][1][t >> 1][r][(t & 1) * 2]) = make_double2(mim[0], mim[1]);
                if (sys.rhs && warp == 0) y_update(X, buf);     // X == p1: the factor warp needs y_{p1} for z(s+1)
            }
            if (sys.rhs && warp == NW - 1 && lane >= 8 && lane < 16)      // recycle the rhs window slot p: global block s+T
                sm.y[p * TS + lane - 8] = sm.ringRhs[s % kRing][lane - 8];
            if (warp == T - 1) {
                // this warp has no M' block: it streams the finished panel s and [A11^{-1} | z] to HBM (coalesced 16-byte stores)
                const double2* src = reinterpret_cast<const double2*>(&sm.raw[buf][0][0][0][0]);
                double2* dst = reinterpret_cast<double2*>(sys.panels + (size_t)s * panel_doubles(T));
                constexpr int NQ = panel_doubles(T) / 2 / 32;      // 16-byte chunks per lane
                constexpr int HQ = (NQ + 1) / 2;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double2 tmp[HQ];
#pragma unroll
                    for (int qq = 0; qq < HQ; ++qq) if (h * HQ + qq < NQ) tmp[qq] = src[lane + 32 * (h * HQ + qq)];
#pragma unroll
                    for (int qq = 0; qq < HQ; ++qq) if (h * HQ + qq < NQ) dst[lane + 32 * (h * HQ + qq)] = tmp[qq];
                }
                cplx* dz = sys.ainvz + (size_t)s * AZ;
                for (int qq = lane; qq < AZ; qq += 32) dz[qq] = sm.ainvz[buf][qq];
            }
            bar_sync(BAR_M, NW * 32);
            // recycle the tile touching slot block p: its panel (raw(s)) is already published, and it now holds
            // entries of global block s+T, whose untouched stencil couplings may already reach the next pivot block.
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                if (i != ip) continue;
                int dI = tI[i] - p1; if (dI < 0) dI += T;      // position inside the window [s+1, s+T]
                int dJ = tJ[i] - p1; if (dJ < 0) dJ += T;
                // row/column distance between the entering block (position T-1) and the other block of the tile
                const int delta = TS * (T - 1 - min(dI, dJ));
                const bool cand = !staged || delta <= TS || (delta + 7 >= nf && delta - 7 <= nf);   // warp-uniform
                cre[i][0] = cre[i][1] = cim[i][0] = cim[i][1] = 0.0;
                if (cand) {
                    int gi = (s + 1 + dI) * TS + g;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        int gj = (s + 1 + dJ) * TS + 2 * t + e;
                        int hi = max(gi, gj), lo = min(gi, gj);    // hi always lies in the entering block s+T
                        cplx v = staged ? prov.get_staged(sm.ringRow[s % kRing], hi, lo) : prov.get(hi, lo);
                        cre[i][e] = v.x;
                        cim[i][e] = v.y;
                    }
                }
            }
            // pass 1: the tile of the next pivot column (and the pivot block after it) first, then publish
            if (s + 1 < S) {
#pragma unroll
                for (int i = 0; i < TPW; ++i) {
                    if (i == ip1) {
                        if (i != ip) update(i, buf);            // the recycled {p,p1} tile is fresh: no update
                        dump(i, p1, buf ^ 1);
                    } else if (i == idg) {
                        if (i != ip) update(i, buf);            // (T == 2: p2 == p, already recycled)
                        dump_diag(i, buf);                      // pivot block of panel s+2, updated through panel s
                    }
                }
                bar_arrive(BAR_RAW, NTHR);
            }
            if (sys.rhs && myX >= 0 && warp != 0) y_update(myX, buf);
            // pass 2: the rest of the trailing window
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                if (i == ip || i == ip1 || i == idg) continue;
                update(i, buf);
            }
        }
    } else {
        // =============================== factor warp ===============================
        const int i = g, j0 = 2 * t;
        cplx a0, a1;                                        // A11^{-1}[i][j0], [i][j0+1] of the CURRENT panel
        bool bad = false;
        // in-register Gauss-Jordan inversion of an 8x8 complex block held as (row i, cols j0, j0+1)
        // Fraction-free Gauss-Jordan (validated in numpy): rows i != k take  row_i <- (p row_i - a_ik row_k) 2^-e
        // with an exact power-of-two rescale, so no reciprocal sits on the 8-pivot dependency chain; every row carries
        // its accumulated scale q_i and the true inverse is  a_ij / q_i, formed with one reciprocal per row at the end.
        auto invert = [&]() {
            cplx q = mk(1.0, 0.0);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int srcRow = 4 * k + t, srcCol = 4 * i + (k >> 1), srcPiv = 4 * k + (k >> 1);
                const cplx mine = (k & 1) ? a1 : a0;
                const cplx pk = mk(__shfl_sync(0xffffffffu, mine.x, srcPiv), __shfl_sync(0xffffffffu, mine.y, srcPiv));
                const cplx f = mk(__shfl_sync(0xffffffffu, mine.x, srcCol), __shfl_sync(0xffffffffu, mine.y, srcCol));
                const cplx ck = mk(__shfl_sync(0xffffffffu, q.x, 4 * k), __shfl_sync(0xffffffffu, q.y, 4 * k));
                cplx r0 = mk(__shfl_sync(0xffffffffu, a0.x, srcRow), __shfl_sync(0xffffffffu, a0.y, srcRow));
                cplx r1 = mk(__shfl_sync(0xffffffffu, a1.x, srcRow), __shfl_sync(0xffffffffu, a1.y, srcRow));
                if (j0 == k) r0 = ck;                       // slot (k,k) switches to the right-hand block: value c_k
                if (j0 + 1 == k) r1 = ck;
                const double mag = fmax(fabs(pk.x), fabs(pk.y));
                if (!(mag > 1e-290) || !(mag < 1e290)) bad = true;
                const int eb = (__double2hiint(mag) >> 20) & 0x7ff;
                const double sc = __hiloint2double((2046 - eb) << 20, 0);          // 2^-(exponent of the pivot), exact
                const cplx ps = mk(pk.x * sc, pk.y * sc), fs = mk(f.x * sc, f.y * sc);
                if (i == k) {
                    a0 = r0; a1 = r1; q = pk;
                } else {
                    const cplx b0 = (j0 == k) ? mk(0.0, 0.0) : a0, b1 = (j0 + 1 == k) ? mk(0.0, 0.0) : a1;
                    cplx n0 = ps * b0, n1 = ps * b1;
                    cfma(n0, -fs, r0);
                    cfma(n1, -fs, r1);
                    a0 = n0; a1 = n1;
                    q = ps * q;
                }
            }
            const double den = fma(q.x, q.x, q.y * q.y);
            if (!(den > 0.0) || isinf(den)) bad = true;
            const double iden = __drcp_rn(den);
            const cplx qi = mk(q.x * iden, -q.y * iden);
            a0 = a0 * qi;
            a1 = a1 * qi;
        };
        auto publish = [&](int buf) {      // -A11^{-1} as B-fragments, plain A11^{-1} for the TMA store
            // B-fragment layout wants plane[kk][n][tt] = -Ainv[4kk+tt][n]; Ainv is symmetric, so write -Ainv[i][j] at [j>>2][i][j&3]
            *reinterpret_cast<double2*>(&sm.nainv[buf][0][j0 >> 2][i][j0 & 3]) = make_double2(-a0.x, -a1.x);
            *reinterpret_cast<double2*>(&sm.nainv[buf][1][j0 >> 2][i][j0 & 3]) = make_double2(-a0.y, -a1.y);
            sm.ainvz[buf][i * 8 + j0] = a0;
            sm.ainvz[buf][i * 8 + j0 + 1] = a1;
        };
        // prologue: panel 0
        bar_sync(BAR_RAW, NTHR);
        {
            double2 lre = *reinterpret_cast<const double2*>(&sm.raw[0][0][j0 >> 2][i][j0 & 3]);
            double2 lim = *reinterpret_cast<const double2*>(&sm.raw[0][1][j0 >> 2][i][j0 & 3]);
            a0 = mk(lre.x, lim.x); a1 = mk(lre.y, lim.y);
            invert();
            publish(0);
        }
        for (int s = 0; s < S; ++s) {
            const int p = s % T, p1 = (s + 1) % T, buf = s & 1;
            const int rp = p * TS;
            if (bad && lane == 0) { sm.fail = 1; if (sys.status) *sys.status = kErrSingular; }
            if (sys.rhs) {
                // z_i = sum_j Ainv[i][j] y_p[j] : two terms per lane, reduced over the 4 lanes of a row
                cplx acc = a0 * sm.y[rp + j0] + a1 * sm.y[rp + j0 + 1];
#pragma unroll
                for (int off = 1; off <= 2; off <<= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
                }
                if (t == 0) sm.ainvz[buf][64 + i] = acc;
            } else if (t == 0) sm.ainvz[buf][64 + i] = mk(0.0, 0.0);
            bar_arrive(BAR_INV, NTHR);                      // panel s may be applied (bar.arrive orders the prior smem writes)
            if (bad) break;
            // ring prefetch for the block entering the window kPre steps from now (loads stay in flight over the inversion)
            double rowv; cplx rhsv;
            ring_fetch(s + kPre + T, lane, rowv, rhsv);
            if (s + 1 < S) {
                // ---- early inversion of the next pivot block:  A11(s+1) = D(s+1; through s-1) + M'_{p1} raw_{p1}^T ----
                double dre[2], dim_[2];
                {
                    double2 vre = *reinterpret_cast<const double2*>(&sm.dnext[buf ^ 1][0][i][j0]);
                    double2 vim = *reinterpret_cast<const double2*>(&sm.dnext[buf ^ 1][1][i][j0]);
                    dre[0] = vre.x; dre[1] = vre.y; dim_[0] = vim.x; dim_[1] = vim.y;
                }
                const int r1 = p1 * TS + g;
                double mre[2] = {0.0, 0.0}, mim[2] = {0.0, 0.0}, mr2[2] = {0.0, 0.0}, mi2[2] = {0.0, 0.0};
                double bre[2], bim[2];
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    bre[kk] = sm.raw[buf][0][kk][r1][t]; bim[kk] = sm.raw[buf][1][kk][r1][t];
                    double nre = sm.nainv[buf][0][kk][g][t], nim = sm.nainv[buf][1][kk][g][t];
                    dmma884(mre, bre[kk], nre);
                    dmma884(mim, bre[kk], nim);
                    dmma884(mr2, -bim[kk], nim);
                    dmma884(mi2, bim[kk], nre);
                }
                mre[0] += mr2[0]; mre[1] += mr2[1]; mim[0] += mi2[0]; mim[1] += mi2[1];
                *reinterpret_cast<double2*>(&sm.mscr[0][t >> 1][g][(t & 1) * 2]) = make_double2(mre[0], mre[1]);
                *reinterpret_cast<double2*>(&sm.mscr[1][t >> 1][g][(t & 1) * 2]) = make_double2(mim[0], mim[1]);
                __syncwarp();
                double dr2[2] = {0.0, 0.0}, di2[2] = {0.0, 0.0};
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    double are = sm.mscr[0][kk][g][t], aim = sm.mscr[1][kk][g][t];
                    dmma884(dre, are, bre[kk]);
                    dmma884(dim_, are, bim[kk]);
                    dmma884(dr2, -aim, bim[kk]);
                    dmma884(di2, aim, bre[kk]);
                }
                a0 = mk(dre[0] + dr2[0], dim_[0] + di2[0]); a1 = mk(dre[1] + dr2[1], dim_[1] + di2[1]);
                invert();
                publish(buf ^ 1);
            }
            // ring slot of step s+kPre: its previous content (step s+kPre-kRing) was consumed before BAR_M(s+kPre-kRing)
            sm.ringRow[(s + kPre) % kRing][lane >> 3][lane & 7] = rowv;
            if (lane < 8) sm.ringRhs[(s + kPre) % kRing][lane] = rhsv;
            if (s + 1 < S) bar_sync(BAR_RAW, NTHR);         // raw(s+1) complete (y_p(s+1) final as well)
        }
    }
    fence_proxy_async_all();                                // the panels are read back below through the async proxy (TMA)
    __threadfence();
    __syncthreads();
    if (!sys.rhs || sm.fail) return;

    // =============================== fused back-substitution ===============================
    //   x_p = z_s - A11_s^{-1} (raw_s^T x_rest),  s = S-1 .. 0   (x window circular in smem: reuse sm.y)
    // The factor is streamed back with TMA bulk loads, kBackStages-1 panels in flight.
    constexpr int NST = FactorSmem<T>::kBackStages;
    for (int i = tid; i < R; i += NTHR) sm.y[i] = mk(0.0, 0.0);
    if (tid == 0) {
        for (int q = 0; q < NST; ++q) mbar_init(&sm.mbar[q], 1);
        fence_mbar_init();
    }
    __syncthreads();
    constexpr uint32_t PBYTES = panel_doubles(T) * 8;
    auto issue = [&](int s) {     // thread 0: prefetch panel s and [A11^{-1}|z]_s into stage (S-1-s) % NST
        int st = (S - 1 - s) % NST;
        mbar_arrive_expect_tx(&sm.mbar[st], PBYTES + AZ * 16);
        bulk_g2s(&sm.stage[st][0][0][0][0], sys.panels + (size_t)s * panel_doubles(T), PBYTES, &sm.mbar[st]);
        bulk_g2s(&sm.stageAZ[st][0], sys.ainvz + (size_t)s * AZ, AZ * 16, &sm.mbar[st]);
    };
    if (tid == 0) {
        fence_proxy_async();
        for (int q = 0; q < NST - 1 && S - 1 - q >= 0; ++q) issue(S - 1 - q);
    }
    constexpr int NRG = NTHR / 8;        // row groups
    for (int s = S - 1; s >= 0; --s) {
        const int p = s % T, it = S - 1 - s, st = it % NST;
        if (tid == 0 && s - (NST - 1) >= 0) issue(s - (NST - 1));      // stage freed at the end of the previous step
        mbar_wait(&sm.mbar[st], (uint32_t)((it / NST) & 1));
        const int c = tid & 7, rg = tid >> 3;
        cplx acc = mk(0.0, 0.0);
        for (int r = rg; r < R; r += NRG) {
            if ((r >> 3) == p) continue;
            cplx rv = mk(sm.stage[st][0][c >> 2][r][c & 3], sm.stage[st][1][c >> 2][r][c & 3]);
            cfma(acc, rv, sm.y[r]);
        }
        // reduce the 4 row groups inside a warp (lanes differing in bits 3,4)
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
        if (lane < 8) sm.part[warp][lane] = acc;
        __syncthreads();
        if (warp == 0) {
            // d_k = sum_w part[w][k] (lanes k + 8j hold partial sums), then x_i = z_i - sum_k Ainv[i][k] d_k
            cplx d = mk(0.0, 0.0);
            for (int w = (lane >> 3); w < NW + 1; w += 4) d += sm.part[w][lane & 7];
#pragma unroll
            for (int off = 8; off <= 16; off <<= 1) {
                d.x += __shfl_xor_sync(0xffffffffu, d.x, off);
                d.y += __shfl_xor_sync(0xffffffffu, d.y, off);
            }
            const int ii = lane >> 2, tt = lane & 3;     // lane handles k = 2tt, 2tt+1 of row ii
            cplx d0 = mk(__shfl_sync(0xffffffffu, d.x, 2 * tt), __shfl_sync(0xffffffffu, d.y, 2 * tt));
            cplx d1 = mk(__shfl_sync(0xffffffffu, d.x, 2 * tt + 1), __shfl_sync(0xffffffffu, d.y, 2 * tt + 1));
            cplx xv = sm.stageAZ[st][ii * 8 + 2 * tt] * d0 + sm.stageAZ[st][ii * 8 + 2 * tt + 1] * d1;
#pragma unroll
            for (int off = 1; off <= 2; off <<= 1) {
                xv.x += __shfl_xor_sync(0xffffffffu, xv.x, off);
                xv.y += __shfl_xor_sync(0xffffffffu, xv.y, off);
            }
            if (tt == 0) {
                cplx xo = sm.stageAZ[st][64 + ii] - xv;
                sm.y[p * TS + ii] = xo;
                int gidx = s * TS + ii;
                if (gidx < N) sys.x[gidx] = xo;
            }
        }
        __syncthreads();
    }
}

}  // namespace hmcmt
