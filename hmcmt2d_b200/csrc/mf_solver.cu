// Host orchestration of the multifrontal solver (see mf_solver.cuh, mf_kernels.cuh).
#include "mf_solver.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "mf_kernels.cuh"

namespace hmcmt {
namespace mf {

namespace {
struct PerDeviceOnceMf {
    bool done[64] = {};
    bool need() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
        if (done[dev]) return false;
        done[dev] = true;
        return true;
    }
};
constexpr size_t kMaxSmem = 227 * 1024;
constexpr int kSolveWarpMaxFp = 64;  // solves: warp-per-front kernels up to this front size
// solves: one warp per front (record-driven kernels) for small fronts with few children, one CTA per front otherwise
inline bool solve_by_warp(const Front& F) { return !F.isBig && F.fp() <= kSolveWarpMaxFp && F.nChild <= kDescChildren; }

template <int NW>
int launch_small(cudaStream_t st, const Tables& tb, const SmallDesc* list, int n, int nsys, size_t smem) {
    static PerDeviceOnceMf once;
    if (once.need()) HMCMT_CUDA_TRY(cudaFuncSetAttribute(mf_small_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    mf_small_kernel<NW><<<dim3(n, nsys), NW * 32, smem, st>>>(tb, list);
    return kOk;
}
}  // namespace

template <typename Tp>
int Solver::upload(const std::vector<Tp>& h, Tp** d) {
    *d = nullptr;
    const size_t n = std::max<size_t>(h.size(), 1);
    if (cudaMalloc(d, n * sizeof(Tp)) != cudaSuccess) return kErrAlloc;
    owned.push_back(*d);
    bytes += n * sizeof(Tp);
    if (!h.empty() && cudaMemcpy(*d, h.data(), h.size() * sizeof(Tp), cudaMemcpyHostToDevice) != cudaSuccess) return kErrCuda;
    return kOk;
}

int Solver::upload_solve_descs(const std::vector<int>& fr, const SolveDesc** d) {
    std::vector<SolveDesc> v;
    v.reserve(fr.size());
    for (int k : fr) {
        const Front& F = S.fronts[k];
        SolveDesc sd{};
        sd.sp = F.sp; sd.up = F.up; sd.s = F.s; sd.u = F.u; sd.cbp = F.cbp; sd.rowPtr = F.rowPtr; sd.updOff = (int)F.updOff;
        sd.nChild = F.nChild;
        sd.gOff = S.chunks[F.chunkPtr].gOff; sd.mOff = S.chunks[F.chunkPtr].mOff;
        for (int c = 0; c < std::min(F.nChild, kDescChildren); ++c) {
            const Front& C = S.fronts[S.children[F.childPtr + c]];
            sd.cRel[c] = C.rowPtr; sd.cUpd[c] = (int)C.updOff; sd.cU[c] = C.u;
        }
        for (int i = 0; i < std::min(F.u, kDescRows); ++i) sd.rows[i] = S.rows[F.rowPtr + i];
        v.push_back(sd);
    }
    SolveDesc* p = nullptr;
    int rc = upload(v, &p);
    *d = p;
    return rc;
}

Solver* Solver::create(Symbolic&& S, int nsys, int maxRhs, int64_t valCount, int* rc) {
    Solver* s = new Solver();
    s->S = std::move(S);
    int r = s->build(nsys, maxRhs, valCount);
    if (rc) *rc = r;
    if (r != kOk) { delete s; return nullptr; }
    return s;
}

Solver::~Solver() {
    if (d_prof) {
        unsigned long long h[48] = {};
        cudaDeviceSynchronize();
        cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
        for (int c = 0; c < 4; ++c) {
            const unsigned long long* q = h + 8 * c;
            if (q[6])
                fprintf(stderr, "[hmcmt_b200] mf_small_kernel<%d> cycles per front (%llu fronts, %.2f pivot blocks each): zero %.0f  orig %.0f  "
                        "children %.0f  mirror %.0f  sweep %.0f  output %.0f\n", 2 << c, q[6], (double)q[7] / q[6], (double)q[0] / q[6],
                        (double)q[1] / q[6], (double)q[2] / q[6], (double)q[3] / q[6], (double)q[4] / q[6], (double)q[5] / q[6]);
            const unsigned long long* w = h + 32 + 4 * c;
            if (q[7])
                fprintf(stderr, "[hmcmt_b200]     factorisation of the front, cycles per pivot block: 8x8 inversions %.0f  rest of the pivot-block sweep %.0f  M' = F21 (-G) %.0f  U += M' F21^T %.0f\n",
                        (double)w[0] / q[7], (double)w[1] / q[7], (double)w[2] / q[7], (double)w[3] / q[7]);
        }
    }
    for (void* p : owned) cudaFree(p);
}

int Solver::build(int nsys_, int maxRhs_, int64_t valCount_) {
    nsys = nsys_; maxRhs = std::max(1, maxRhs_); valCount = valCount_;
    int rc;
#define MF_TRY(x) do { rc = (x); if (rc) return rc; } while (0)
    MF_TRY(upload(S.fronts, &d_fronts));
    MF_TRY(upload(S.rows, &d_rows));
    MF_TRY(upload(S.rel, &d_rel));
    MF_TRY(upload(S.children, &d_children));
    MF_TRY(upload(S.orig, &d_orig));
    MF_TRY(upload(S.chunks, &d_chunks));
    MF_TRY(upload(S.pos2orig, &d_pos2orig));
    auto dalloc = [&](void** p, size_t nbytes) {
        *p = nullptr;
        if (nbytes == 0) nbytes = 16;
        if (cudaMalloc(p, nbytes) != cudaSuccess) return (int)kErrAlloc;
        owned.push_back(*p);
        bytes += nbytes;
        return (int)kOk;
    };
    MF_TRY(dalloc((void**)&d_fac, (size_t)nsys * S.factorDoubles * sizeof(double)));
    for (int p = 0; p < 2; ++p) MF_TRY(dalloc((void**)&d_arena[p], (size_t)nsys * S.arenaDoubles[p] * sizeof(double)));
    MF_TRY(dalloc((void**)&d_vals, (size_t)nsys * valCount * sizeof(cplx)));
    MF_TRY(dalloc((void**)&d_v, (size_t)nsys * maxRhs * S.Np * sizeof(cplx)));
    MF_TRY(dalloc((void**)&d_upd, (size_t)nsys * maxRhs * S.updEntries * sizeof(cplx)));
    if (const char* e = std::getenv("HMCMT_MF_PROF"); e && std::atoi(e)) {
        MF_TRY(dalloc((void**)&d_prof, 48 * sizeof(unsigned long long)));
        cudaMemset(d_prof, 0, 48 * sizeof(unsigned long long));
    }

    // launch schedule
    std::vector<int> invMapsHost;
    sched.assign(S.maxDepth + 1, DepthSchedule());
    for (int d = 0; d <= S.maxDepth; ++d) {
        DepthSchedule& D = sched[d];
        const std::vector<int>& sm = S.byDepthSmall[d];
        const std::vector<int>& bg = S.byDepthBig[d];
        int* p = nullptr;
        D.nSmall = (int)sm.size();
        if (D.nSmall) {
            int mx = 0;
            D.smallSmem = 0;
            std::vector<SmallDesc> descs;
            for (int k : sm) {
                const Front& F = S.fronts[k];
                mx = std::max(mx, F.fp());
                D.smallSmem = std::max(D.smallSmem, mf_front_smem_bytes(F.fp() / 8, F.sp / 8, mf_rel_total(S, k)));
                SmallDesc sd{};
                sd.sp = F.sp; sd.up = F.up; sd.s = F.s; sd.nChild = F.nChild; sd.nOrig = F.nOrig; sd.origPtr = F.origPtr;
                sd.par = F.depth & 1; sd.front = k;
                sd.gOff = S.chunks[F.chunkPtr].gOff; sd.mOff = S.chunks[F.chunkPtr].mOff; sd.uOff = F.frontOff;
                for (int c = 0; c < std::min(F.nChild, kDescChildren); ++c) {
                    const Front& C = S.fronts[S.children[F.childPtr + c]];
                    sd.cOff[c] = C.frontOff; sd.cLd[c] = C.ldU(); sd.cFirst[c] = C.offU(); sd.cU[c] = C.u; sd.cRel[c] = C.rowPtr;
                }
                descs.push_back(sd);
            }
            SmallDesc* pd = nullptr;
            MF_TRY(upload(descs, &pd));
            D.smallDescs = pd;
            auto knob = [](const char* name, int dflt) { const char* e = std::getenv(name); return e ? std::atoi(e) : dflt; };
            // warps per front by the largest front of the launch (measured optimum at cfg2; one warp per front needs no CTA barrier)
            const int f1 = knob("HMCMT_MF_FP1", 40), f2 = knob("HMCMT_MF_FP2", 40), f4 = knob("HMCMT_MF_FP4", 64), f8 = knob("HMCMT_MF_FP8", 96);
            D.smallWarps = mx <= f1 ? 1 : (mx <= f2 ? 2 : (mx <= f4 ? 4 : (mx <= f8 ? 8 : 16)));
        }
        D.nBig = (int)bg.size();
        D.bigBytes = (size_t)S.bigDoublesAtDepth[d] * sizeof(double);
        // solves: one warp per front up to kSolveWarpMaxFp rows, one CTA per front above
        std::vector<int> sw, sc(bg);
        for (int k : sm) (solve_by_warp(S.fronts[k]) ? sw : sc).push_back(k);
        D.nSolveWarp = (int)sw.size();
        D.nSolveCta = (int)sc.size();
        if (D.nSolveWarp) MF_TRY(upload_solve_descs(sw, &D.solveWarpList));
        if (D.nSolveCta) { MF_TRY(upload(sc, &p)); D.solveCtaList = p; }
        {
            int mxc = 0;
            for (int k : sc) mxc = std::max(mxc, S.fronts[k].fp());
            const char* e = std::getenv("HMCMT_MF_SOLVE_THREADS");      // 0: by front size
            const int forced = e ? std::atoi(e) : 0;
            D.solveCtaThreads = forced ? forced : (mxc <= 144 ? 128 : kSolveMfThreads);
        }
        if (!D.nBig) continue;
        MF_TRY(upload(bg, &p));
        D.bigList = p;
        std::vector<int2> op;
        int maxChunk = 0;
        for (int k : bg) {
            const Front& F = S.fronts[k];
            for (int e = 0; e < F.nOrig; ++e) op.push_back(make_int2(k, F.origPtr + e));
            maxChunk = std::max(maxChunk, F.nChunk);
        }
        int2* p2 = nullptr;
        D.nOrigPairs = (int)op.size();
        if (D.nOrigPairs) { MF_TRY(upload(op, &p2)); D.origPairs = p2; }
        {
            std::vector<AsmTile> tiles;
            for (int k : bg) {
                const Front& F = S.fronts[k];
                const int fp = F.fp(), nt = (fp + 63) / 64;
                const int invPtr = (int)invMapsHost.size();
                for (int c = 0; c < F.nChild; ++c) {
                    const Front& C = S.fronts[S.children[F.childPtr + c]];
                    std::vector<int> inv(fp, -1);
                    for (int i = 0; i < C.u; ++i) inv[S.rel[C.rowPtr + i]] = i;
                    invMapsHost.insert(invMapsHost.end(), inv.begin(), inv.end());
                }
                for (int bi = 0; bi < nt; ++bi)
                    for (int bj = 0; bj <= bi; ++bj) tiles.push_back(AsmTile{k, bi, bj, invPtr});
            }
            AsmTile* pt = nullptr;
            MF_TRY(upload(tiles, &pt));
            D.asmTiles = pt;
            D.nAsmTiles = (int)tiles.size();
        }
        for (int c = 0; c < maxChunk; ++c) {
            DepthSchedule::ChunkStep cs;
            std::vector<int> inv;
            std::vector<GemmJob> jobs;
            std::vector<GemmTile> pt, st;
            int maxSc = 0;
            for (int k : bg) {
                const Front& F = S.fronts[k];
                if (F.nChunk <= c) continue;
                const Chunk& ch = S.chunks[F.chunkPtr + c];
                const int sc = ch.p1 - ch.p0, fp = F.fp(), mr = fp - ch.p1, par = F.depth & 1;
                inv.push_back(k);
                maxSc = std::max(maxSc, sc);
                if (mr <= 0) continue;
                GemmJob jp{};      // M = F21 G
                jp.aOff = F.frontOff; jp.aSp = par; jp.ldA = fp; jp.rA = ch.p1; jp.kA = ch.p0;
                jp.bOff = ch.gOff; jp.bSp = 2; jp.ldB = sc; jp.rB = 0; jp.kB = 0;
                jp.cOff = ch.mOff; jp.cSp = 2; jp.ldC = mr; jp.rC = 0; jp.cC = 0;
                jp.m = mr; jp.n = sc; jp.K = sc; jp.lower = 0; jp.beta = 0; jp.alpha = 1.0;
                const int jpi = (int)jobs.size();
                jobs.push_back(jp);
                for (int bi = 0; bi < (mr + 63) / 64; ++bi)
                    for (int bj = 0; bj < (sc + 63) / 64; ++bj) pt.push_back(GemmTile{jpi, bi, bj});
                GemmJob js{};      // F22 -= M F21^T (lower tiles)
                js.aOff = ch.mOff; js.aSp = 2; js.ldA = mr; js.rA = 0; js.kA = 0;
                js.bOff = F.frontOff; js.bSp = par; js.ldB = fp; js.rB = ch.p1; js.kB = ch.p0;
                js.cOff = F.frontOff; js.cSp = par; js.ldC = fp; js.rC = ch.p1; js.cC = ch.p1;
                js.m = mr; js.n = mr; js.K = sc; js.lower = 1; js.beta = 1; js.alpha = -1.0;
                const int jsi = (int)jobs.size();
                jobs.push_back(js);
                for (int bi = 0; bi < (mr + 63) / 64; ++bi)
                    for (int bj = 0; bj <= bi; ++bj) st.push_back(GemmTile{jsi, bi, bj});
            }
            MF_TRY(upload(inv, &p));
            cs.invList = p; cs.nInv = (int)inv.size();
            cs.invSmem = mf_front_smem_bytes(maxSc / 8, maxSc / 8);
            GemmJob* pj = nullptr;
            GemmTile* ptile = nullptr;
            MF_TRY(upload(jobs, &pj));
            cs.jobs = pj;
            MF_TRY(upload(pt, &ptile));
            cs.panelTiles = ptile; cs.nPanelTiles = (int)pt.size();
            MF_TRY(upload(st, &ptile));
            cs.schurTiles = ptile; cs.nSchurTiles = (int)st.size();
            D.chunkSteps.push_back(cs);
        }
    }
    MF_TRY(upload(invMapsHost, &d_invMaps));
    solveSmem = (size_t)(2 * S.maxFp + 16) * sizeof(cplx);
    if (solveSmem > kMaxSmem) return kErrArg;
#undef MF_TRY
    return kOk;
}

int Solver::set_mt_values(cudaStream_t st, int N, const MtValSys* dSys, int sys0, int n) {
    if (valCount < 3 * (int64_t)N) return kErrArg;
    if (n < 0) n = nsys - sys0;
    if (sys0 < 0 || n < 1 || sys0 + n > nsys) return kErrArg;
    mf_mt_vals_kernel<<<dim3((3 * N + 255) / 256, n), 256, 0, st>>>(N, dSys + sys0, d_vals + (size_t)sys0 * valCount, valCount);
    HMCMT_CUDA_TRY(cudaGetLastError());
    return kOk;
}

int Solver::factor(cudaStream_t st, int* dStatus, int64_t* nLaunches, int sys0, int n) {
    if (n < 0) n = this->nsys - sys0;
    if (sys0 < 0 || n < 1 || sys0 + n > this->nsys) return kErrArg;
    const int nsys = n;      // systems of this call: the kernels index them 0..n-1 from the offset bases below
    Tables tb{};
    tb.fronts = d_fronts; tb.rows = d_rows; tb.rel = d_rel; tb.children = d_children; tb.orig = d_orig; tb.chunks = d_chunks;
    tb.pos2orig = d_pos2orig; tb.fac = d_fac + (size_t)sys0 * S.factorDoubles;
    tb.arena[0] = d_arena[0] + (size_t)sys0 * S.arenaDoubles[0]; tb.arena[1] = d_arena[1] + (size_t)sys0 * S.arenaDoubles[1];
    tb.facStride = S.factorDoubles; tb.arenaStride[0] = S.arenaDoubles[0]; tb.arenaStride[1] = S.arenaDoubles[1];
    tb.vals = d_vals + (size_t)sys0 * valCount; tb.valStride = valCount; tb.status = dStatus ? dStatus + sys0 : nullptr; tb.prof = d_prof;
    static PerDeviceOnceMf once;
    if (once.need()) {
        HMCMT_CUDA_TRY(cudaFuncSetAttribute(mf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
        HMCMT_CUDA_TRY(cudaFuncSetAttribute(mf_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes));
    }
    int64_t nl = 0;
    for (int d = S.maxDepth; d >= 0; --d) {
        const DepthSchedule& D = sched[d];
        if (D.nSmall) {
            int rc = kOk;
            if (D.smallWarps == 1) rc = launch_small<1>(st, tb, D.smallDescs, D.nSmall, nsys, D.smallSmem);
            else if (D.smallWarps == 2) rc = launch_small<2>(st, tb, D.smallDescs, D.nSmall, nsys, D.smallSmem);
            else if (D.smallWarps == 4) rc = launch_small<4>(st, tb, D.smallDescs, D.nSmall, nsys, D.smallSmem);
            else if (D.smallWarps == 8) rc = launch_small<8>(st, tb, D.smallDescs, D.nSmall, nsys, D.smallSmem);
            else rc = launch_small<16>(st, tb, D.smallDescs, D.nSmall, nsys, D.smallSmem);
            if (rc) return rc;
            ++nl;
        }
        if (!D.nBig) continue;
        if (D.nAsmTiles) {
            mf_asm_gather_kernel<<<dim3(D.nAsmTiles, nsys), 256, 0, st>>>(tb, (const AsmTile*)D.asmTiles, d_invMaps);
            ++nl;
        }
        if (D.nOrigPairs) {
            mf_asm_orig_kernel<<<dim3((D.nOrigPairs + 255) / 256, nsys), 256, 0, st>>>(tb, D.origPairs, D.nOrigPairs);
            ++nl;
        }
        for (size_t c = 0; c < D.chunkSteps.size(); ++c) {
            const DepthSchedule::ChunkStep& cs = D.chunkSteps[c];
            if (!cs.nInv) continue;
            mf_inv_kernel<<<dim3(cs.nInv, nsys), kInvWarps * 32, cs.invSmem, st>>>(tb, cs.invList, (int)c);
            ++nl;
            if (cs.nPanelTiles) {
                mf_gemm_kernel<<<dim3(cs.nPanelTiles, nsys), kGemmThreads, kGemmSmemBytes, st>>>(
                    tb, (const GemmJob*)cs.jobs, (const GemmTile*)cs.panelTiles);
                ++nl;
            }
            if (cs.nSchurTiles) {
                mf_gemm_kernel<<<dim3(cs.nSchurTiles, nsys), kGemmThreads, kGemmSmemBytes, st>>>(
                    tb, (const GemmJob*)cs.jobs, (const GemmTile*)cs.schurTiles);
                ++nl;
            }
        }
    }
    HMCMT_CUDA_TRY(cudaGetLastError());
    if (nLaunches) *nLaunches += nl;
    return kOk;
}

int Solver::add_rhs_pattern(const std::vector<unsigned char>& nz) {
    if ((int)nz.size() != S.N) return kErrArg;
    std::vector<char> active(S.K, 0);
    for (int k = 0; k < S.K; ++k) {          // elimination order: children before parents
        const Front& F = S.fronts[k];
        for (int i = 0; i < F.s && !active[k]; ++i) {
            const int o = S.pos2orig[F.cbp + i];
            if (o >= 0 && nz[o]) active[k] = 1;
        }
        if (active[k] && F.parent >= 0) active[F.parent] = 1;
    }
    FwdLists L;
    int rc;
#define MF_TRY(x) do { rc = (x); if (rc) return rc; } while (0)
    for (int d = 0; d <= S.maxDepth; ++d) {
        std::vector<int> sw, sc;
        for (int k : S.byDepthBig[d]) if (active[k]) sc.push_back(k);
        for (int k : S.byDepthSmall[d]) if (active[k]) (solve_by_warp(S.fronts[k]) ? sw : sc).push_back(k);
        const SolveDesc* pd = nullptr;
        if (!sw.empty()) MF_TRY(upload_solve_descs(sw, &pd));
        L.warpList.push_back(pd);
        L.nWarp.push_back((int)sw.size());
        int* p = nullptr;
        if (!sc.empty()) MF_TRY(upload(sc, &p));
        L.ctaList.push_back(sc.empty() ? nullptr : p);
        L.nCta.push_back((int)sc.size());
        L.nFronts += (int)(sw.size() + sc.size());
    }
#undef MF_TRY
    patterns.push_back(std::move(L));
    return (int)patterns.size();
}

int Solver::fwd_fronts(int pattern) const {
    return pattern >= 1 && pattern <= (int)patterns.size() ? patterns[pattern - 1].nFronts : S.K;
}

int Solver::solve(cudaStream_t st, int nrhs, const cplx* B, int64_t ldb, cplx* X, int64_t ldx, int64_t* nLaunches, int sys0, int n,
                  int pattern) {
    if (nrhs < 1 || nrhs > maxRhs || pattern < 0 || pattern > (int)patterns.size()) return kErrArg;
    if (n < 0) n = this->nsys - sys0;
    if (sys0 < 0 || n < 1 || sys0 + n > this->nsys) return kErrArg;
    const int nsys = n;
    Tables tb{};
    tb.fronts = d_fronts; tb.rows = d_rows; tb.rel = d_rel; tb.children = d_children; tb.orig = d_orig; tb.chunks = d_chunks;
    tb.pos2orig = d_pos2orig; tb.fac = d_fac + (size_t)sys0 * S.factorDoubles;
    tb.arena[0] = d_arena[0] + (size_t)sys0 * S.arenaDoubles[0]; tb.arena[1] = d_arena[1] + (size_t)sys0 * S.arenaDoubles[1];
    tb.facStride = S.factorDoubles; tb.arenaStride[0] = S.arenaDoubles[0]; tb.arenaStride[1] = S.arenaDoubles[1];
    tb.vals = d_vals; tb.valStride = valCount; tb.status = nullptr; tb.prof = nullptr;
    const size_t vec0 = (size_t)sys0 * nrhs;
    // the workspaces are laid out for maxRhs vectors per system
    SolveArgs sa{B + vec0 * ldb, X + vec0 * ldx, ldb, ldx, d_v + (size_t)sys0 * maxRhs * S.Np, d_upd + (size_t)sys0 * maxRhs * S.updEntries,
                 (int64_t)S.Np, S.updEntries, nrhs};
    static PerDeviceOnceMf once;
    if (once.need()) {
        HMCMT_CUDA_TRY(cudaFuncSetAttribute(mf_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
        HMCMT_CUDA_TRY(cudaFuncSetAttribute(mf_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    }
    const int nvec = nsys * nrhs;
    int64_t nl = 0;
    const FwdLists* P = pattern ? &patterns[pattern - 1] : nullptr;
    if (P) {
        // fronts outside the pattern are not visited: their pivot parts and update vectors are zero
        HMCMT_CUDA_TRY(cudaMemsetAsync(sa.v, 0, (size_t)nvec * S.Np * sizeof(cplx), st));
        HMCMT_CUDA_TRY(cudaMemsetAsync(sa.upd, 0, (size_t)nvec * S.updEntries * sizeof(cplx), st));
    }
    for (int d = S.maxDepth; d >= 0; --d) {
        const DepthSchedule& D = sched[d];
        const int nW = P ? P->nWarp[d] : D.nSolveWarp, nC = P ? P->nCta[d] : D.nSolveCta;
        const SolveDesc* wl = P ? P->warpList[d] : D.solveWarpList;
        const int* cl = P ? P->ctaList[d] : D.solveCtaList;
        if (nW) {
            mf_fwd_warp_kernel<<<dim3((nW + kSolveWarpsPerCta - 1) / kSolveWarpsPerCta, nvec), kSolveWarpsPerCta * 32, 0, st>>>(tb, sa, wl, nW);
            ++nl;
        }
        if (nC) {
            mf_fwd_kernel<<<dim3(nC, nvec), D.solveCtaThreads, solveSmem, st>>>(tb, sa, cl);
            ++nl;
        }
    }
    for (int d = 0; d <= S.maxDepth; ++d) {
        const DepthSchedule& D = sched[d];
        if (D.nSolveWarp) {
            mf_bwd_warp_kernel<<<dim3((D.nSolveWarp + kSolveWarpsPerCta - 1) / kSolveWarpsPerCta, nvec), kSolveWarpsPerCta * 32, 0, st>>>(
                tb, sa, D.solveWarpList, D.nSolveWarp);
            ++nl;
        }
        if (D.nSolveCta) {
            mf_bwd_kernel<<<dim3(D.nSolveCta, nvec), D.solveCtaThreads, solveSmem, st>>>(tb, sa, D.solveCtaList);
            ++nl;
        }
    }
    HMCMT_CUDA_TRY(cudaGetLastError());
    if (nLaunches) *nLaunches += nl;
    return kOk;
}

}  // namespace mf
}  // namespace hmcmt
