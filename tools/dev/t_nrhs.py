"""Level-1 solves with several right-hand sides: time per call vs nrhs (is the factor re-read per vector?).  3-D div-grad of the
reference's own test size (MUMPS/test/testDivGrad.jl:9) and the cfg2 stencil system."""
import os, sys, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hmcmt2d_b200 import lib


def ddx(n):
    return sp.diags([-np.ones(n), np.ones(n)], [0, 1], shape=(n, n + 1))


def divgrad(n1, n2, n3):
    I = sp.identity
    Div = sp.hstack([sp.kron(I(n3), sp.kron(I(n2), ddx(n1))), sp.kron(I(n3), sp.kron(ddx(n2), I(n1))), sp.kron(ddx(n3), sp.kron(I(n2), I(n1)))])
    return (Div @ Div.T).tocsc()


def stencil(nl, nf, rng):
    N = nl * nf
    d = 4 + rng.random(N) + 1j * rng.random(N)
    e1, e2 = -rng.random(N), -rng.random(N)
    e1[np.arange(N) % nf == 0] = 0
    return sp.diags([d, e1[1:], e1[1:], e2[nf:], e2[nf:]], [0, -1, 1, -nf, nf], format="csc")


rng = np.random.default_rng(0)
os.environ["HMCMT_SHIM_SOLVER"] = "mf"
for name, A in [("divgrad 32x32x16", (divgrad(32, 32, 16) + 1j * sp.identity(32 * 32 * 16) * 0.1).tocsc()), ("stencil 199x99", stencil(199, 99, rng))]:
    n = A.shape[0]
    F = lib.factorMUMPS(A, 2)
    for nrhs in (1, 2, 4, 8, 10, 16):
        b = rng.standard_normal((n, nrhs)) + 1j * rng.standard_normal((n, nrhs))
        x = lib.applyMUMPS(F, b)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); x = lib.applyMUMPS(F, b); ts.append(time.perf_counter() - t0)
        res = max(np.linalg.norm(A @ x[:, i] - b[:, i]) / np.linalg.norm(b[:, i]) for i in range(nrhs))
        print(f"[nrhs] {name}: nrhs {nrhs:2d}: {min(ts) * 1e3:8.3f} ms per call ({min(ts) * 1e3 / nrhs:7.3f} per vector)  residual {res:.1e}", flush=True)
    lib.destroyMUMPS(F)
