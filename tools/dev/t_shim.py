import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, '/root/repo')
from hmcmt2d_b200 import lib
print(lib.load().hmcmt_version())
rng = np.random.default_rng(0)
def grid_matrix(nl, nf, cplx=True):
    N = nl*nf
    d = 4 + rng.random(N) + (1j*rng.random(N) if cplx else 0)
    e1 = -rng.random(N); e2 = -rng.random(N)
    f = np.arange(N) % nf
    e1[f == 0] = 0
    A = sp.diags([d, e1[1:], e1[1:], e2[nf:], e2[nf:]], [0, -1, 1, -nf, nf], format='csc')
    return A
for (nl, nf, cplx) in [(7, 10, True), (13, 55, True), (9, 51, False), (20, 99, True), (5, 104, True), (40, 30, True)]:
    A = grid_matrix(nl, nf, cplx)
    N = A.shape[0]
    rhs = rng.standard_normal((N, 3)) + (1j*rng.standard_normal((N, 3)) if cplx else 0)
    t = time.time()
    x = lib.solveMUMPS(A, rhs, 1)
    dt = time.time() - t
    res = np.linalg.norm(A @ x - rhs, axis=0) / np.linalg.norm(rhs, axis=0)
    xr = spla.splu(A).solve(rhs)
    print(nl, nf, cplx, 'resid', res.max(), 'vs splu', np.abs(x - xr).max() / np.abs(xr).max(), 'time', round(dt, 3), x.dtype)
