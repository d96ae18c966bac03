// CPU check of the multifrontal symbolic tables (hmcmt2d_b200/csrc/mf_symbolic.h): executes the numeric phase the CUDA kernels
// perform — same tables, same depth-by-depth schedule, same front arithmetic (G = F11^-1, M = F21 G, U = F22 - M F21^T,
// chunked pivots) — in plain std::complex on the host and verifies the residual of A x = b.  Test infrastructure only.
//   usage: mf_tables_check grid <nl> <nf> <leaf> <fSmall>   |   mf_tables_check graph3d <nx> <ny> <nz> <leaf> <fSmall>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>

#include "../../hmcmt2d_b200/csrc/mf_symbolic.h"

using namespace hmcmt::mf;
typedef std::complex<double> cd;

static void invert(std::vector<cd>& A, int n) {      // Gauss-Jordan without pivoting
    std::vector<cd> B((size_t)n * n, cd(0));
    for (int i = 0; i < n; ++i) B[(size_t)i * n + i] = 1.0;
    for (int k = 0; k < n; ++k) {
        cd p = 1.0 / A[(size_t)k * n + k];
        for (int j = 0; j < n; ++j) { A[(size_t)k * n + j] *= p; B[(size_t)k * n + j] *= p; }
        for (int i = 0; i < n; ++i) if (i != k) {
            cd f = A[(size_t)i * n + k];
            if (f == cd(0)) continue;
            for (int j = 0; j < n; ++j) { A[(size_t)i * n + j] -= f * A[(size_t)k * n + j]; B[(size_t)i * n + j] -= f * B[(size_t)k * n + j]; }
        }
    }
    A = B;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::vector<Entry> ent;
    std::vector<std::vector<int>> sn;
    int N = 0, fSmall = 144;
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> un(0.1, 1.0);
    if (!strcmp(argv[1], "grid")) {
        int nl = atoi(argv[2]), nf = atoi(argv[3]), leaf = atoi(argv[4]);
        fSmall = atoi(argv[5]);
        N = nl * nf;
        mf_grid_entries(nl, nf, ent);
        const int cross = argc > 6 ? atoi(argv[6]) : 0;      // four-way cross separators for boxes up to this size
        const int push = argc > 7 ? atoi(argv[7]) : 0;        // leaf excess over a multiple of 8 handed to the separator above
        mf_order_grid(nl, nf, leaf, sn, cross, push);
    } else {
        int nx = atoi(argv[2]), ny = atoi(argv[3]), nz = atoi(argv[4]), leaf = atoi(argv[5]);
        fSmall = atoi(argv[6]);
        N = nx * ny * nz;
        auto id = [&](int i, int j, int k) { return (k * ny + j) * nx + i; };
        int src = 0;
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
            int q = id(i, j, k);
            ent.push_back(Entry{q, q, src++});
            if (i > 0) ent.push_back(Entry{q, id(i - 1, j, k), src++});
            if (j > 0) ent.push_back(Entry{q, id(i, j - 1, k), src++});
            if (k > 0) ent.push_back(Entry{q, id(i, j, k - 1), src++});
        }
        std::vector<int> ptr(N + 1, 0), adj;
        for (auto& e : ent) if (e.row != e.col) { ++ptr[e.row + 1]; ++ptr[e.col + 1]; }
        for (int i = 0; i < N; ++i) ptr[i + 1] += ptr[i];
        adj.resize(ptr[N]);
        std::vector<int> at(ptr.begin(), ptr.end() - 1);
        for (auto& e : ent) if (e.row != e.col) { adj[at[e.row]++] = e.col; adj[at[e.col]++] = e.row; }
        mf_order_graph(N, ptr, adj, leaf, sn);
    }
    // values: diagonally dominant complex symmetric
    int nsrc = 0;
    for (auto& e : ent) nsrc = std::max(nsrc, e.src + 1);
    std::vector<cd> vals(nsrc);
    std::vector<double> rowsum(N, 0.0);
    for (auto& e : ent) if (e.row != e.col) { double v = -un(rng); vals[e.src] = v; rowsum[e.row] += -v; rowsum[e.col] += -v; }
    for (auto& e : ent) if (e.row == e.col) vals[e.src] = cd(rowsum[e.row] + 0.01, 0.3 * un(rng));
    Symbolic S;
    if (!mf_symbolic(N, sn, ent, fSmall, S)) { printf("symbolic failed\n"); return 1; }
    int nbig = 0;
    for (auto& F : S.fronts) nbig += F.isBig;
    printf("N %d Np %d K %d depth %d maxFp %d big %d factor MB %.1f arena MB %.1f/%.1f flops %.3e\n", S.N, S.Np, S.K, S.maxDepth, S.maxFp, nbig,
           S.factorDoubles * 8 / 1e6, S.arenaDoubles[0] * 8 / 1e6, S.arenaDoubles[1] * 8 / 1e6, S.flops);
    if (getenv("MF_STATS")) {
        // per depth: fronts handled by the single-CTA kernel / the large-front path, pivot blocks, largest front, modelled flops
        for (int d = S.maxDepth; d >= 0; --d) {
            int ns = 0, nb = 0, pb = 0, mx = 0;
            double fl = 0.0;
            for (auto& F : S.fronts) if (F.depth == d) {
                (F.isBig ? nb : ns)++; pb += F.sp / 8; mx = std::max(mx, F.fp());
                const double a = F.sp, b = F.up;
                fl += 8.0 * (0.5 * a * a * a + a * a * b + 0.5 * a * b * b);
            }
            printf("  depth %2d: small %5d big %3d pivot blocks %5d max fp %3d flops %.2e\n", d, ns, nb, pb, mx, fl);
        }
        if (getenv("MF_STATS")[0] == '2') return 0;
    }
    // checks of the schedule invariants
    for (int k = 0; k < S.K; ++k) {
        const Front& F = S.fronts[k];
        for (int c = 0; c < F.nChild; ++c) {
            const Front& C = S.fronts[S.children[F.childPtr + c]];
            if (C.depth != F.depth + 1 || C.parent != k) { printf("depth/parent invariant broken\n"); return 1; }
        }
        if (F.sp % 8 || F.up % 8) { printf("padding broken\n"); return 1; }
    }
    // numeric phase, depth by depth
    std::vector<std::vector<cd>> U(S.K), G(S.chunks.size()), M(S.chunks.size());
    auto run_front = [&](int k) {
        const Front& F = S.fronts[k];
        const int fp = F.fp();
        std::vector<cd> A((size_t)fp * fp, cd(0));
        for (int e = 0; e < F.nOrig; ++e) {
            const OrigEntry& oe = S.orig[F.origPtr + e];
            cd v = oe.src < 0 ? cd(1.0) : vals[oe.src];
            if (oe.lrow < oe.lcol) { printf("orig entry above the diagonal\n"); exit(1); }
            A[(size_t)oe.lrow * fp + oe.lcol] += v;
        }
        for (int c = 0; c < F.nChild; ++c) {
            const int ck = S.children[F.childPtr + c];
            const Front& C = S.fronts[ck];
            const int* rel = S.rel.data() + C.rowPtr;
            for (int i = 0; i < C.u; ++i) for (int j = 0; j <= i; ++j) {
                if (rel[i] < rel[j] || rel[i] < 0 || rel[i] >= fp) { printf("rel not monotone\n"); exit(1); }
                A[(size_t)rel[i] * fp + rel[j]] += U[ck][(size_t)i * C.up + j];
            }
            std::vector<cd>().swap(U[ck]);
        }
        for (int i = 0; i < fp; ++i) for (int j = 0; j < i; ++j) A[(size_t)j * fp + i] = A[(size_t)i * fp + j];
        for (int c = 0; c < F.nChunk; ++c) {
            const Chunk& ch = S.chunks[F.chunkPtr + c];
            const int sc = ch.p1 - ch.p0, mr = fp - ch.p1;
            std::vector<cd> g((size_t)sc * sc);
            for (int i = 0; i < sc; ++i) for (int j = 0; j < sc; ++j) g[(size_t)i * sc + j] = A[(size_t)(ch.p0 + i) * fp + ch.p0 + j];
            invert(g, sc);
            std::vector<cd> m((size_t)mr * sc, cd(0));
            for (int i = 0; i < mr; ++i) for (int kk = 0; kk < sc; ++kk) {
                cd a = A[(size_t)(ch.p1 + i) * fp + ch.p0 + kk];
                if (a == cd(0)) continue;
                for (int j = 0; j < sc; ++j) m[(size_t)i * sc + j] += a * g[(size_t)kk * sc + j];
            }
            for (int i = 0; i < mr; ++i) for (int j = 0; j < mr; ++j) {
                cd acc = 0;
                for (int kk = 0; kk < sc; ++kk) acc += m[(size_t)i * sc + kk] * A[(size_t)(ch.p1 + j) * fp + ch.p0 + kk];
                A[(size_t)(ch.p1 + i) * fp + ch.p1 + j] -= acc;
            }
            G[F.chunkPtr + c] = std::move(g);
            M[F.chunkPtr + c] = std::move(m);
        }
        U[k].assign((size_t)F.up * F.up, cd(0));
        for (int i = 0; i < F.up; ++i) for (int j = 0; j < F.up; ++j) U[k][(size_t)i * F.up + j] = A[(size_t)(F.sp + i) * fp + F.sp + j];
    };
    for (int d = S.maxDepth; d >= 0; --d) {
        for (int k : S.byDepthBig[d]) run_front(k);
        for (int k : S.byDepthSmall[d]) run_front(k);
    }
    // solve
    std::vector<cd> b(N), x(N), v(S.Np, cd(0)), upd(S.updEntries, cd(0));
    for (auto& z : b) z = cd(un(rng) - 0.5, un(rng) - 0.5);
    for (int d = S.maxDepth; d >= 0; --d)
        for (int pass = 0; pass < 2; ++pass)
            for (int k : (pass ? S.byDepthSmall[d] : S.byDepthBig[d])) {
                const Front& F = S.fronts[k];
                const int fp = F.fp();
                std::vector<cd> w(fp, cd(0));
                for (int i = 0; i < F.s; ++i) w[i] = b[S.pos2orig[F.cbp + i]];
                for (int c = 0; c < F.nChild; ++c) {
                    const Front& C = S.fronts[S.children[F.childPtr + c]];
                    for (int i = 0; i < C.u; ++i) w[S.rel[C.rowPtr + i]] += upd[C.updOff + i];
                }
                for (int c = 0; c < F.nChunk; ++c) {
                    const Chunk& ch = S.chunks[F.chunkPtr + c];
                    const int sc = ch.p1 - ch.p0, mr = fp - ch.p1;
                    for (int i = 0; i < mr; ++i) {
                        cd acc = 0;
                        for (int kk = 0; kk < sc; ++kk) acc += M[F.chunkPtr + c][(size_t)i * sc + kk] * w[ch.p0 + kk];
                        w[ch.p1 + i] -= acc;
                    }
                }
                for (int i = 0; i < F.sp; ++i) v[F.cbp + i] = w[i];
                for (int i = 0; i < F.up; ++i) upd[F.updOff + i] = w[F.sp + i];
            }
    for (int d = 0; d <= S.maxDepth; ++d)
        for (int pass = 0; pass < 2; ++pass)
            for (int k : (pass ? S.byDepthSmall[d] : S.byDepthBig[d])) {
                const Front& F = S.fronts[k];
                const int fp = F.fp();
                std::vector<cd> xf(fp, cd(0));
                for (int i = 0; i < F.sp; ++i) xf[i] = v[F.cbp + i];
                for (int i = 0; i < F.u; ++i) xf[F.sp + i] = v[S.rows[F.rowPtr + i]];
                for (int c = F.nChunk - 1; c >= 0; --c) {
                    const Chunk& ch = S.chunks[F.chunkPtr + c];
                    const int sc = ch.p1 - ch.p0, mr = fp - ch.p1;
                    std::vector<cd> t(sc, cd(0));
                    for (int kk = 0; kk < sc; ++kk) {
                        cd acc = 0;
                        for (int j = 0; j < sc; ++j) acc += G[F.chunkPtr + c][(size_t)j * sc + kk] * xf[ch.p0 + j];
                        for (int i = 0; i < mr; ++i) acc -= M[F.chunkPtr + c][(size_t)i * sc + kk] * xf[ch.p1 + i];
                        t[kk] = acc;
                    }
                    for (int kk = 0; kk < sc; ++kk) xf[ch.p0 + kk] = t[kk];
                }
                for (int i = 0; i < F.sp; ++i) { v[F.cbp + i] = xf[i]; if (i < F.s) x[S.pos2orig[F.cbp + i]] = xf[i]; }
            }
    // residual
    std::vector<cd> r(b);
    for (auto& e : ent) {
        r[e.row] -= vals[e.src] * x[e.col];
        if (e.row != e.col) r[e.col] -= vals[e.src] * x[e.row];
    }
    double nr = 0, nb = 0;
    for (int i = 0; i < N; ++i) { nr += std::norm(r[i]); nb += std::norm(b[i]); }
    double rel = std::sqrt(nr / nb);
    printf("relative residual %.3e\n", rel);
    return rel < 1e-10 ? 0 : 1;
}
