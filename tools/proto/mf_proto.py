"""Algorithm prototype of the nested-dissection multifrontal solver (csrc/mf_*.cuh): numpy restatement of the
arithmetic the CUDA kernels perform, checked against SciPy SuperLU on the MT systems.  Not product code.

    python tools/proto/mf_proto.py 200 100      (ny nz)

Front arithmetic (pivot-free, complex symmetric, no conjugation), front = [pivots s | update rows u]:
    G = F11^{-1}      by block Gauss-Jordan with 8x8 pivot blocks (explicit 8x8 inverses, as gj_invert8)
    M = F21 G
    U = F22 - M F21^T
solve:  forward  w2 -= M w1 ;  backward  x1 = G w1 - M^T x2
Separators longer than SMAX columns are eliminated in chunks of <= SMAX columns on the same front.
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

MU0 = 4e-7 * np.pi
SMAX = 96
LEAF = 32          # leaf boxes hold at most this many nodes


def planes(ylen, zlen, sig, mode):
    """5-point stencil of SURVEY.md A.3 on the interior nodes; returns dict of 2-D arrays indexed [kn-1][jn-1]."""
    ny, nz = len(ylen), len(zlen)
    s = sig.reshape(nz, ny)
    v = np.full_like(s, 1.0 / MU0) if mode == 0 else 1.0 / s
    w = s if mode == 0 else np.full_like(s, MU0)
    dy, dz = ylen, zlen
    # edges in y between node (jn,kn) and (jn+1,kn): jn = 0..ny-1, kn = 1..nz-1
    wy = 0.5 * (dz[:-1, None] * v[:-1, :] + dz[1:, None] * v[1:, :]) / dy[None, :]          # [kn-1][jc]
    wz = 0.5 * (dy[None, :-1] * v[:, :-1] + dy[None, 1:] * v[:, 1:]) / dz[:, None]          # [kc][jn-1]
    a = dy[None, :] * dz[:, None] * w
    mass = 0.25 * (a[:-1, :-1] + a[:-1, 1:] + a[1:, :-1] + a[1:, 1:])
    diag = wy[:, :-1] + wy[:, 1:] + wz[:-1, :] + wz[1:, :]
    return dict(diag=diag, mass=mass, wy=wy, wz=wz)


def system(ylen, zlen, sig, mode, freq):
    """Sparse A (internal ordering: fast axis = shorter one) and the grid shape (nl, nf)."""
    P = planes(ylen, zlen, sig, mode)
    n2, n1 = P["diag"].shape         # n2 = nz-1 (k), n1 = ny-1 (j)
    om = 2 * np.pi * freq
    D = P["diag"] + 1j * om * P["mass"]
    cy = -P["wy"][:, 1:-1]            # coupling (jn,kn)-(jn+1,kn), jn=1..n1-1 : [k][j]
    cz = -P["wz"][1:-1, :]            # coupling (jn,kn)-(jn,kn+1): [k][j]
    fastZ = n2 <= n1
    if fastZ:                          # q = (jn-1)*n2 + (kn-1)
        idx = (np.arange(n1)[None, :] * n2 + np.arange(n2)[:, None])
        nl, nf = n1, n2
    else:
        idx = (np.arange(n2)[:, None] * n1 + np.arange(n1)[None, :])
        nl, nf = n2, n1
    N = n1 * n2
    rows = [idx.ravel(), idx[:, :-1].ravel(), idx[:, 1:].ravel(), idx[:-1, :].ravel(), idx[1:, :].ravel()]
    cols = [idx.ravel(), idx[:, 1:].ravel(), idx[:, :-1].ravel(), idx[1:, :].ravel(), idx[:-1, :].ravel()]
    vals = [D.ravel(), cy.ravel(), cy.ravel(), cz.ravel(), cz.ravel()]
    A = sp.csc_matrix((np.concatenate(vals).astype(complex), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    return A, nl, nf


# ---------------------------------------------------------------------------------------------------------
# symbolic: geometric nested dissection of the nl x nf grid (q = l*nf + f)

def nd_order(nl, nf, leaf=LEAF):
    """-> list of supernodes (arrays of internal indices) in elimination (post) order."""
    out = []

    def rec(l0, l1, f0, f1):
        nL, nF = l1 - l0, f1 - f0
        if nL <= 0 or nF <= 0:
            return
        if nL * nF <= leaf:
            out.append((np.arange(l0, l1)[:, None] * nf + np.arange(f0, f1)[None, :]).ravel())
            return
        if nL >= nF:
            mid = (l0 + l1) // 2
            rec(l0, mid, f0, f1)
            rec(mid + 1, l1, f0, f1)
            out.append(mid * nf + np.arange(f0, f1))
        else:
            mid = (f0 + f1) // 2
            rec(l0, l1, f0, mid)
            rec(l0, l1, mid + 1, f1)
            out.append(np.arange(l0, l1) * nf + mid)
    rec(0, nl, 0, nf)
    return out


def symbolic(A, snodes):
    """perm, per-supernode (col range, update rows in permuted numbering, parent)."""
    N = A.shape[0]
    perm = np.concatenate(snodes)
    assert len(perm) == N and len(np.unique(perm)) == N
    inv = np.empty(N, dtype=np.int64)
    inv[perm] = np.arange(N)
    Ap = A[perm][:, perm].tocsc()
    K = len(snodes)
    cb = np.concatenate([[0], np.cumsum([len(s) for s in snodes])])
    sn_of = np.repeat(np.arange(K), np.diff(cb))
    pending = [[] for _ in range(K)]
    rows, parent = [None] * K, np.full(K, -1)
    for k in range(K):
        c0, c1 = cb[k], cb[k + 1]
        r = Ap.indices[Ap.indptr[c0]:Ap.indptr[c1]]
        parts = [r[r >= c1]] + pending[k]
        st = np.unique(np.concatenate(parts)) if parts else np.zeros(0, np.int64)
        rows[k] = st
        if len(st):
            p = sn_of[st[0]]
            parent[k] = p
            pending[p].append(st[st >= cb[p + 1]])
        pending[k] = None
    return perm, Ap, cb, rows, parent


def block_gj_inverse(F):
    """in-place style block Gauss-Jordan with 8x8 pivot blocks, no pivoting (what mf_inv does)."""
    n = F.shape[0]
    A = F.copy()
    for k0 in range(0, n, 8):
        k1 = min(k0 + 8, n)
        P = np.linalg.inv(A[k0:k1, k0:k1])
        rest = np.r_[0:k0, k1:n]
        C = A[np.ix_(rest, range(k0, k1))] @ P                    # A_ik P
        R = A[np.ix_(range(k0, k1), rest)].copy()                 # A_kj
        A[np.ix_(rest, rest)] -= C @ R
        A[np.ix_(range(k0, k1), rest)] = P @ R
        A[np.ix_(rest, range(k0, k1))] = -C
        A[k0:k1, k0:k1] = P
        # sign convention: after the step rows/cols k hold [P R ; -C]; standard GJ: row k <- P*row k ; col k <- -C
    # The plain (unsymmetric-storage) Gauss-Jordan inverse: fix the sign bookkeeping by recomputing from definition
    return A


def gj_inverse_sym(F):
    """Symmetric sweep with 8x8 blocks: after sweeping all blocks the array holds -F^{-1}; returns F^{-1}."""
    n = F.shape[0]
    A = F.copy()
    for k0 in range(0, n, 8):
        k1 = min(k0 + 8, n)
        kk = np.arange(k0, k1)
        rest = np.r_[0:k0, k1:n]
        P = np.linalg.inv(A[np.ix_(kk, kk)])
        C = A[np.ix_(rest, kk)] @ P                               # A_ik P
        A[np.ix_(rest, rest)] -= C @ A[np.ix_(kk, rest)]
        A[np.ix_(rest, kk)] = C
        A[np.ix_(kk, rest)] = C.T
        A[np.ix_(kk, kk)] = -P
    return -A


class MF:
    def __init__(self, A, nl, nf, leaf=LEAF, smax=SMAX):
        t0 = time.time()
        self.N = A.shape[0]
        self.snodes = nd_order(nl, nf, leaf)
        self.perm, self.Ap, self.cb, self.rows, self.parent = symbolic(A, self.snodes)
        self.K = len(self.snodes)
        self.smax = smax
        self.children = [[] for _ in range(self.K)]
        for k in range(self.K):
            if self.parent[k] >= 0:
                self.children[self.parent[k]].append(k)
        self.t_sym = time.time() - t0

    def stats(self):
        s = np.diff(self.cb)
        u = np.array([len(r) for r in self.rows])
        f = s + u
        macs = (s ** 3 / 2 + s ** 2 * u + s * u ** 2 / 2).sum()          # sweep count (explicit inverse)
        macs_ldl = (s ** 3 / 6 + s ** 2 * u / 2 + s * u ** 2 / 2).sum()
        fac = (s * f).sum()
        depth = np.zeros(self.K, int)
        for k in range(self.K - 1, -1, -1):
            if self.parent[k] >= 0:
                depth[k] = depth[self.parent[k]] + 1
        return dict(K=self.K, maxf=int(f.max()), maxs=int(s.max()), flops_sweep=8 * macs, flops_ldl=8 * macs_ldl,
                    factor_entries=int(fac), depth=int(depth.max()), frontmem=int((f ** 2).sum()),
                    by_depth=[(d, int((depth == d).sum()), int(s[depth == d].max()), int(u[depth == d].max()),
                               float(8 * (s ** 3 / 2 + s ** 2 * u + s * u ** 2 / 2)[depth == d].sum()))
                              for d in range(depth.max() + 1)])

    def factor(self):
        K, cb, rows = self.K, self.cb, self.rows
        Ap = self.Ap
        self.G, self.M, self.chunks = [None] * K, [None] * K, [None] * K
        U = [None] * K
        for k in range(K):
            c0, c1 = cb[k], cb[k + 1]
            s, r = c1 - c0, rows[k]
            idx = np.concatenate([np.arange(c0, c1), r])
            f = len(idx)
            F = Ap[idx][:, c0:c1].toarray()
            Ff = np.zeros((f, f), complex)
            Ff[:, :s] = F
            Ff[:s, s:] = F[s:, :].T
            pos = {}
            for c in self.children[k]:
                rel = np.searchsorted(idx, rows[c])
                assert np.all(idx[rel] == rows[c])
                Ff[np.ix_(rel, rel)] += U[c]
                U[c] = None
            # eliminate the pivots in chunks of <= smax columns on the same front
            Gs, Ms, ch = [], [], []
            p0 = 0
            while p0 < s:
                p1 = min(p0 + self.smax, s)
                G = gj_inverse_sym(Ff[p0:p1, p0:p1])
                M = Ff[p1:, p0:p1] @ G
                Ff[p1:, p1:] -= M @ Ff[p1:, p0:p1].T
                Gs.append(G); Ms.append(M); ch.append((p0, p1))
                p0 = p1
            self.G[k], self.M[k], self.chunks[k] = Gs, Ms, ch
            U[k] = Ff[s:, s:]

    def solve(self, b):
        K, cb, rows = self.K, self.cb, self.rows
        y = b[self.perm].astype(complex)
        upd = [None] * K
        W = [None] * K
        for k in range(K):
            c0, c1 = cb[k], cb[k + 1]
            idx = np.concatenate([np.arange(c0, c1), rows[k]])
            w = np.zeros(len(idx), complex)
            w[:c1 - c0] = y[c0:c1]
            for c in self.children[k]:
                rel = np.searchsorted(idx, rows[c])
                w[rel] += upd[c]
                upd[c] = None
            for (p0, p1), M in zip(self.chunks[k], self.M[k]):
                w[p1:] -= M @ w[p0:p1]
            W[k] = w[:c1 - c0].copy()
            upd[k] = w[c1 - c0:]
        x = np.zeros(self.N, complex)
        for k in range(K - 1, -1, -1):
            c0, c1 = cb[k], cb[k + 1]
            s = c1 - c0
            xf = np.concatenate([np.zeros(s, complex), x[rows[k]]])
            for (p0, p1), G, M in reversed(list(zip(self.chunks[k], self.G[k], self.M[k]))):
                xf[p0:p1] = G @ W[k][p0:p1] - M.T @ xf[p1:]
            x[c0:c1] = xf[:s]
        out = np.empty(self.N, complex)
        out[self.perm] = x
        return out


def synthetic(ny, nz, seed=1):
    sys.path.insert(0, ".")
    from hmcmt2d_b200 import synthetic as syn
    mesh = syn.make_mesh(ny, nz)
    nair = 7
    rng = np.random.default_rng(seed)
    sig = np.asarray(mesh.sigma).copy()
    sig[ny * nair:] = np.exp(np.log(0.01) + 0.7 * rng.standard_normal(ny * (nz - nair)))
    return np.asarray(mesh.yLen), np.asarray(mesh.zLen), sig


if __name__ == "__main__":
    ny, nz = int(sys.argv[1]), int(sys.argv[2])
    leaf = int(sys.argv[3]) if len(sys.argv) > 3 else LEAF
    ylen, zlen, sig = synthetic(ny, nz)
    first = True
    for mode in (0, 1):
        for freq in (100.0, 0.3, 0.001):
            A, nl, nf = system(ylen, zlen, sig, mode, freq)
            mf = MF(A, nl, nf, leaf)
            if first:
                st = mf.stats()
                print({k: v for k, v in st.items() if k != "by_depth"}, "symbolic %.2fs" % mf.t_sym)
                for row in st["by_depth"]:
                    print("  depth %2d  fronts %6d  max s %4d  max u %4d  flops %.3e" % row)
                print("band flops 4Nb^2 = %.3e" % (4.0 * A.shape[0] * nf * nf))
                first = False
            rng = np.random.default_rng(5)
            b = rng.standard_normal(A.shape[0]) + 1j * rng.standard_normal(A.shape[0])
            t0 = time.time()
            mf.factor()
            x = mf.solve(b)
            t1 = time.time()
            lu = spla.splu(A)
            xr = lu.solve(b)
            res = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
            resr = np.linalg.norm(A @ xr - b) / np.linalg.norm(b)
            err = np.abs(x - xr).max() / np.abs(xr).max()
            print("mode %d freq %-7g  mf %.1fs  resid %.2e (superlu %.2e)  max|x-x_lu|/max|x| %.2e" % (mode, freq, t1 - t0, res, resr, err))
