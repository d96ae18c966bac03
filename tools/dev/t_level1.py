"""Level-1 call pattern of the unmodified reference on a cfg2-size matrix: factor / solve / solve / destroy per frequency and mode,
the same sparsity pattern every time.  ms per factorisation once the symbolic analysis and the solver pool are warm."""
import os, sys, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hmcmt2d_b200 import lib

rng = np.random.default_rng(0)
nl, nf = 199, 99
N = nl * nf
e1, e2 = -rng.random(N), -rng.random(N)
e1[np.arange(N) % nf == 0] = 0
os.environ["HMCMT_SHIM_SOLVER"] = "mf"


def matrix(w):
    d = 4 + rng.random(N) + 1j * w * rng.random(N)
    return sp.diags([d, e1[1:], e1[1:], e2[nf:], e2[nf:]], [0, -1, 1, -nf, nf], format="csc")


b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
for rep in range(3):
    tf = ts = 0.0
    live = []
    for k in range(10):                       # ten live factorisations, as one per frequency in the reference
        A = matrix(1.0 + k)
        t0 = time.perf_counter()
        F = lib.factorMUMPS(A, 1)
        t1 = time.perf_counter()
        x = lib.applyMUMPS(F, b)
        t2 = time.perf_counter()
        tf += t1 - t0
        ts += t2 - t1
        live.append((F, A, x))
    res = max(np.linalg.norm(A @ x - b) / np.linalg.norm(b) for _, A, x in live)
    for F, _, _ in live:
        lib.destroyMUMPS(F)
    print(f"[level1] round {rep}: factor {tf / 10 * 1e3:.2f} ms, solve {ts / 10 * 1e3:.2f} ms per call, worst residual {res:.1e}", flush=True)
