"""numpy emulation of the circular 8x8-tile window block elimination used by the CUDA
band factorisation kernel (slot logic, recycle/reinit, block forward/backward solves)."""
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla

TS = 8
def run(N, nf, seed=0):
    rng = np.random.default_rng(seed)
    b = nf
    T = (b + 7) // 8 + 1
    R = TS * T
    # grid-like complex symmetric matrix in internal ordering
    dr = 4 + rng.random(N); dm = rng.random(N)
    e1 = -rng.random(N); e2 = -rng.random(N)
    f = np.arange(N) % nf
    e1[f == 0] = 0; e2[:nf] = 0
    def entry(g, c):
        if g < c: g, c = c, g
        if g >= N: return 1.0 if g == c else 0.0
        if g == c: return dr[g] + 1j * dm[g]
        if c == g - 1 and (g % nf) != 0: return e1[g]
        if c == g - nf: return e2[g]
        return 0.0
    A = sp.lil_matrix((N, N), dtype=complex)
    for g in range(N):
        A[g, g] = entry(g, g)
        if g % nf: A[g, g-1] = e1[g]; A[g-1, g] = e1[g]
        if g >= nf: A[g, g-nf] = e2[g]; A[g-nf, g] = e2[g]
    rhs = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    xref = spla.spsolve(A.tocsc(), rhs)
    S = (N + TS - 1) // TS
    # window: slot blocks 0..T-1 ; W[I][J] 8x8 for all (full square here; kernel keeps I<=J)
    W = np.zeros((R, R), dtype=complex)
    glob = np.zeros(T, dtype=int)            # global block number held by each slot block
    for X in range(T): glob[X] = X
    def fill(X):    # (re)initialise every window entry involving slot block X
        for Y in range(T):
            for rr in range(TS):
                for cc in range(TS):
                    v = entry(glob[X]*TS+rr, glob[Y]*TS+cc)
                    W[X*TS+rr, Y*TS+cc] = v; W[Y*TS+cc, X*TS+rr] = v
    for X in range(T): fill(X)
    y = np.zeros(R, dtype=complex)
    for X in range(T):
        for rr in range(TS):
            g = glob[X]*TS+rr; y[X*TS+rr] = rhs[g] if g < N else 0
    raws = np.zeros((S, R, TS), dtype=complex); ainvs = np.zeros((S, TS, TS), dtype=complex); zs = np.zeros((S, TS), dtype=complex)
    for s in range(S):
        p = s % T
        ps = slice(p*TS, p*TS+TS)
        raw = W[:, ps].copy(); A11 = raw[ps, :].copy(); raw[ps, :] = 0
        Ainv = np.linalg.inv(A11)
        W -= raw @ Ainv @ raw.T                    # trailing update (rows/cols of p untouched as raw[p]=0)
        z = Ainv @ y[ps]; y -= raw @ z
        raws[s] = raw; ainvs[s] = Ainv; zs[s] = z
        glob[p] = s + T; W[ps, :] = 0; W[:, ps] = 0; fill(p)
        for rr in range(TS):
            g = glob[p]*TS+rr; y[p*TS+rr] = rhs[g] if g < N else 0
    # backward
    x = np.zeros(R, dtype=complex); xs = np.zeros(S*TS, dtype=complex)
    for s in range(S-1, -1, -1):
        p = s % T; ps = slice(p*TS, p*TS+TS)
        x[ps] = 0
        xp = zs[s] - ainvs[s] @ (raws[s].T @ x)
        x[ps] = xp; xs[s*TS:(s+1)*TS] = xp
    err = np.abs(xs[:N] - xref).max() / np.abs(xref).max()
    return err
for N, nf in [(99*20, 99), (55*13+3, 55), (51*9, 51), (10*7, 10), (104*5, 104)]:
    print(N, nf, run(N, nf))
