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
//   * the "factor warp" inverts the 8x8 pivot block (Gauss-Jordan, no pivoting), forward-
//     eliminates the fused right-hand side and streams the panel to HBM with TMA bulk stores;
//   * the tile warps form M' = raw * (-A11^{-1}) and apply the rank-8 update with
//     mma.sync.m8n8k4.f64 (DMMA.8x8x4, measured full-rate 37.1 TF/s on B200), updating the
//     next pivot column first so the factor warp works one panel ahead (look-ahead);
//   * recycled slots are refilled from the 5-point stencil planes (assembly is fused: the
//     matrix itself is never written to memory).
// The stored factor is { raw_s (8T x 8), A11_s^{-1} (8x8), z_s } per macro-step; it serves the
// forward solve (back-substitution fused below) and the adjoint solve (band_solve.cuh).
#pragma once
#include "common.cuh"

namespace hmcmt {

constexpr int TS = 8;   // tile size == DMMA m,n

// Per-system description for a batched launch (one CTA per system).
struct BandSys {
    // stencil provider (internal ordering, see mt_kernels.cu): A[g][g] = dr[g] + i*omega*dm[g],
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
    cplx* ainv;          // [S][64]
    cplx* z;             // [S][8]  (A11^{-1} * forward-eliminated rhs) — only if rhs != nullptr
    cplx* x;             // [N] solution of the fused system — only if rhs != nullptr
    int* status;         // 0 ok, -10 zero/NaN pivot block
};

__host__ __device__ constexpr int band_T_for(int b) { return (b + 7) / 8 + 1; }
__host__ __device__ constexpr int panel_doubles(int T) { return 16 * TS * T; }

template <int T>
struct FactorCfg {
    static constexpr int NT = T * (T + 1) / 2;
    static constexpr int NW = (T >= 14) ? 15 : (T >= 12) ? 13 : (T >= 10) ? 11 : (T >= 8) ? 12 : (T >= 6) ? 7 : (T >= 4) ? 5 : 3;
    static constexpr int TPW = (NT + NW - 1) / NW;
    static constexpr int R = TS * T;
    static constexpr int NTHREADS = (NW + 1) * 32;
};

template <int T>
struct FactorSmem {
    static constexpr int R = TS * T;
    // [buf][re/im][kk][r][t]  — raw[buf] is one contiguous panel image (16*R doubles)
    double raw[2][2][2][R][4];
    double m[2][2][2][R][4];
    double nainv[2][2][2][8][4];   // -A11^{-1} in B-fragment layout [buf][re/im][kk][n][t]
    cplx ainv[2][64];              // plain A11^{-1} (row-major) for the fused rhs / global store
    cplx zv[2][8];
    cplx y[R];                     // circular window of the forward-eliminated rhs
    cplx gj[64];                   // Gauss-Jordan scratch
    cplx part[16][8];              // back-substitution partial sums
    cplx dots[8];
    unsigned char tI[FactorCfg<T>::NT], tJ[FactorCfg<T>::NT];
    uint64_t mbar[2];
    int fail;
};

enum { BAR_RAW = 1, BAR_INV = 2, BAR_M = 3 };

struct EntryProvider {
    const double *dr, *dm, *e1, *e2;
    const cplx* band;
    double omega;
    int N, nf, b;
    __device__ __forceinline__ cplx get(int hi, int lo) const {
        // hi >= lo
        int d = hi - lo;
        if (hi >= N) return mk(d == 0 ? 1.0 : 0.0, 0.0);      // identity padding past the end
        if (band) return d <= b ? band[(size_t)hi * (b + 1) + d] : mk(0.0, 0.0);
        if (d == 0) return mk(dr[hi], omega * dm[hi]);
        if (d == 1) return mk(e1[hi], 0.0);
        if (d == nf) return mk(e2[hi], 0.0);
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

    // tile table: unordered slot pairs (I <= J)
    for (int i = tid; i < NT; i += NTHR) {
        int I = 0, rem = i;
        while (rem >= T - I) { rem -= T - I; ++I; }
        sm.tI[i] = (unsigned char)I;
        sm.tJ[i] = (unsigned char)(I + rem);
    }
    for (int i = tid; i < R; i += NTHR) {
        cplx v = mk(0.0, 0.0);
        if (sys.rhs && i < N) v = sys.rhs[i];
        sm.y[i] = v;
    }
    if (tid == 0) sm.fail = 0;
    __syncthreads();

    if (warp < NW) {
        // =============================== tile warps ===============================
        double cre[TPW][2], cim[TPW][2];
        int tI[TPW], tJ[TPW];
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
            int tt = warp + NW * i;
            bool valid = tt < NT;
            tI[i] = valid ? sm.tI[tt] : -1;
            tJ[i] = valid ? sm.tJ[tt] : -1;
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
        auto update = [&](int i, int buf) {
            const int ra = tI[i] * TS + g, rb = tJ[i] * TS + g;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                double are = sm.m[buf][0][kk][ra][t], aim = sm.m[buf][1][kk][ra][t];
                double bre = sm.raw[buf][0][kk][rb][t], bim = sm.raw[buf][1][kk][rb][t];
                dmma884(cre[i], are, bre);
                dmma884(cre[i], -aim, bim);
                dmma884(cim[i], are, bim);
                dmma884(cim[i], aim, bre);
            }
        };
#pragma unroll
        for (int i = 0; i < TPW; ++i)
            if (tI[i] == 0) dump(i, 0, 0);      // tI==0 covers every tile touching block 0 (I<=J)
        fence_proxy_async();                    // raw[] is later read by TMA bulk stores (async proxy)
        bar_arrive(BAR_RAW, NTHR);

        for (int s = 0; s < S; ++s) {
            const int p = s % T, p1 = (s + 1) % T, buf = s & 1;
            bar_sync(BAR_INV, NTHR);                       // -A11^{-1}(s) is in smem
            if (sm.fail) break;
            // M'_X = raw_X * (-A11^{-1}) for every slot block X != p
            for (int xi = warp; xi < T - 1; xi += NW) {
                int X = p + 1 + xi; if (X >= T) X -= T;
                double mre[2] = {0.0, 0.0}, mim[2] = {0.0, 0.0};
                const int r = X * TS + g;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    double are = sm.raw[buf][0][kk][r][t], aim = sm.raw[buf][1][kk][r][t];
                    double bre = sm.nainv[buf][0][kk][g][t], bim = sm.nainv[buf][1][kk][g][t];
                    dmma884(mre, are, bre);
                    dmma884(mre, -aim, bim);
                    dmma884(mim, are, bim);
                    dmma884(mim, aim, bre);
                }
                *reinterpret_cast<double2*>(&sm.m[buf][0][t >> 1][r][(t & 1) * 2]) = make_double2(mre[0], mre[1]);
                *reinterpret_cast<double2*>(&sm.m[buf][1][t >> 1][r][(t & 1) * 2]) = make_double2(mim[0], mim[1]);
            }
            bar_sync(BAR_M, NW * 32);
            // recycle slot block p first: its panel (raw(s)) is already published, and it now holds
            // global block s+T whose untouched stencil entries may already couple to the next pivot block.
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                if (tI[i] < 0 || (tI[i] != p && tJ[i] != p)) continue;
                int dI = tI[i] - p1; if (dI < 0) dI += T;      // position inside the window [s+1, s+T]
                int dJ = tJ[i] - p1; if (dJ < 0) dJ += T;
                int gi = (s + 1 + dI) * TS + g;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int gj = (s + 1 + dJ) * TS + 2 * t + e;
                    cplx v = prov.get(max(gi, gj), min(gi, gj));
                    cre[i][e] = v.x;
                    cim[i][e] = v.y;
                }
            }
            // pass 1: tiles of the next pivot column first, then publish them (look-ahead)
            if (s + 1 < S) {
#pragma unroll
                for (int i = 0; i < TPW; ++i) {
                    if (tI[i] < 0 || (tI[i] != p1 && tJ[i] != p1)) continue;
                    if (tI[i] != p && tJ[i] != p) update(i, buf);       // the recycled {p,p1} tile is fresh: no update
                    dump(i, p1, buf ^ 1);
                }
                fence_proxy_async();
                bar_arrive(BAR_RAW, NTHR);
            }
            // pass 2: the rest of the trailing window
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                if (tI[i] < 0 || tI[i] == p || tJ[i] == p || tI[i] == p1 || tJ[i] == p1) continue;
                update(i, buf);
            }
        }
    } else {
        // =============================== factor warp ===============================
        const int i = g, j0 = 2 * t;
        for (int s = 0; s < S; ++s) {
            const int p = s % T, buf = s & 1;
            bar_sync(BAR_RAW, NTHR);                       // raw(s) complete
            // ---- 8x8 Gauss-Jordan inversion of the pivot block (pivot-free) ----
            const int rp = p * TS;
            cplx a0 = mk(sm.raw[buf][0][j0 >> 2][rp + i][j0 & 3], sm.raw[buf][1][j0 >> 2][rp + i][j0 & 3]);
            cplx a1 = mk(sm.raw[buf][0][(j0 + 1) >> 2][rp + i][(j0 + 1) & 3], sm.raw[buf][1][(j0 + 1) >> 2][rp + i][(j0 + 1) & 3]);
            sm.gj[i * 8 + j0] = a0;
            sm.gj[i * 8 + j0 + 1] = a1;
            __syncwarp();
            bool bad = false;
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {
                cplx pk = sm.gj[k * 8 + k];
                cplx f = sm.gj[i * 8 + k];
                cplx r0 = sm.gj[k * 8 + j0], r1 = sm.gj[k * 8 + j0 + 1];
                double mag = cabs2(pk);
                if (!(mag > 0.0) || isinf(mag)) bad = true;
                cplx rinv = crecip(pk);
                r0 = r0 * rinv;
                r1 = r1 * rinv;
                __syncwarp();
                if (i == k) {
                    a0 = (j0 == k) ? rinv : r0;
                    a1 = (j0 + 1 == k) ? rinv : r1;
                } else {
                    a0 = (j0 == k) ? -(f * rinv) : a0 - f * r0;
                    a1 = (j0 + 1 == k) ? -(f * rinv) : a1 - f * r1;
                }
                sm.gj[i * 8 + j0] = a0;
                sm.gj[i * 8 + j0 + 1] = a1;
                __syncwarp();
            }
            if (bad && lane == 0) { sm.fail = 1; if (sys.status) *sys.status = kErrSingular; }
            // lane holds Ainv[i][j0], Ainv[i][j0+1].  B-fragment layout wants plane[kk][n][tt] = -Ainv[4kk+tt][n];
            // Ainv is symmetric, so write -Ainv[i][j] at [kk=j>>2][n=i][tt=j&3].
            if (s > 0) bulk_wait_read0();                  // raw/ainv staging of panel s-2 drained (smem reuse)
            *reinterpret_cast<double2*>(&sm.nainv[buf][0][j0 >> 2][i][j0 & 3]) = make_double2(-a0.x, -a1.x);
            *reinterpret_cast<double2*>(&sm.nainv[buf][1][j0 >> 2][i][j0 & 3]) = make_double2(-a0.y, -a1.y);
            sm.ainv[buf][i * 8 + j0] = a0;
            sm.ainv[buf][i * 8 + j0 + 1] = a1;
            __threadfence_block();
            bar_arrive(BAR_INV, NTHR);
            if (bad) break;
            __syncwarp();
            // ---- fused forward elimination of the rhs:  z = A11^{-1} y_p ;  y_rest -= raw z ----
            if (sys.rhs) {
                if (lane < 8) {
                    cplx acc = mk(0.0, 0.0);
#pragma unroll
                    for (int k = 0; k < 8; ++k) cfma(acc, sm.ainv[buf][lane * 8 + k], sm.y[rp + k]);
                    sm.zv[buf][lane] = acc;
                }
                __syncwarp();
                cplx zl[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) zl[k] = sm.zv[buf][k];
                for (int r = lane; r < R; r += 32) {
                    if ((r >> 3) == p) continue;
                    cplx acc = sm.y[r];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        cplx rv = mk(sm.raw[buf][0][k >> 2][r][k & 3], sm.raw[buf][1][k >> 2][r][k & 3]);
                        cfma(acc, -rv, zl[k]);
                    }
                    sm.y[r] = acc;
                }
                __syncwarp();
                if (lane < 8) {      // recycle the rhs window slot: global block s+T
                    int gnew = (s + T) * TS + lane;
                    sm.y[rp + lane] = (gnew < N) ? sys.rhs[gnew] : mk(0.0, 0.0);
                    sys.z[(size_t)s * 8 + lane] = sm.zv[buf][lane];
                }
            }
            // ---- stream the panel image and A11^{-1} to HBM (TMA bulk stores) ----
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(sys.panels + (size_t)s * panel_doubles(T), &sm.raw[buf][0][0][0][0], panel_doubles(T) * 8);
                bulk_s2g(sys.ainv + (size_t)s * 64, &sm.ainv[buf][0], 64 * 16);
                bulk_commit();
            }
        }
        if (lane == 0) bulk_wait0();
    }
    __threadfence();
    __syncthreads();
    if (!sys.rhs || sm.fail) return;

    // =============================== fused back-substitution ===============================
    //   x_p = z_s - A11_s^{-1} (raw_s^T x_rest),  s = S-1 .. 0   (x window circular in smem: reuse sm.y)
    for (int i = tid; i < R; i += NTHR) sm.y[i] = mk(0.0, 0.0);
    if (tid == 0) {
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    constexpr uint32_t PBYTES = panel_doubles(T) * 8;
    auto issue = [&](int s) {     // thread 0: prefetch panel s and A11^{-1}_s into buffer s&1
        int bf = s & 1;
        mbar_arrive_expect_tx(&sm.mbar[bf], PBYTES + 64 * 16);
        bulk_g2s(&sm.raw[bf][0][0][0][0], sys.panels + (size_t)s * panel_doubles(T), PBYTES, &sm.mbar[bf]);
        bulk_g2s(&sm.ainv[bf][0], sys.ainv + (size_t)s * 64, 64 * 16, &sm.mbar[bf]);
    };
    if (tid == 0) { fence_proxy_async(); issue(S - 1); }
    uint32_t phase[2] = {0, 0};
    constexpr int NRG = NTHR / 8;        // row groups
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
        // reduce the 4 row groups inside a warp (lanes differing in bits 3,4)
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
        if (lane < 8) sm.part[warp][lane] = acc;
        __syncthreads();
        if (tid < 8) {
            cplx d = mk(0.0, 0.0);
            for (int w = 0; w < NW + 1; ++w) d += sm.part[w][tid];
            sm.dots[tid] = d;
        }
        __syncthreads();
        if (tid < 8) {
            cplx xv = sys.z[(size_t)s * 8 + tid];
#pragma unroll
            for (int k = 0; k < 8; ++k) cfma(xv, -sm.ainv[bf][tid * 8 + k], sm.dots[k]);
            sm.y[p * TS + tid] = xv;
            int gidx = s * TS + tid;
            if (gidx < N) sys.x[gidx] = xv;
        }
        __syncthreads();
    }
}

}  // namespace hmcmt
