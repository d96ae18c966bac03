import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, scipy.sparse as sp
from hmcmt2d_b200 import lib
os.environ["HMCMT_SHIM_SOLVER"] = "mf"
rng = np.random.default_rng(0)
nl, nf = 199, 99
N = nl * nf
d = 4 + rng.random(N) + 1j * rng.random(N)
e1, e2 = -rng.random(N), -rng.random(N)
e1[np.arange(N) % nf == 0] = 0
A = sp.diags([d, e1[1:], e1[1:], e2[nf:], e2[nf:]], [0, -1, 1, -nf, nf], format="csc")
F = lib.factorMUMPS(A, 2)
nrhs = int(sys.argv[1])
b = rng.standard_normal((N, nrhs)) + 1j * rng.standard_normal((N, nrhs))
x = lib.applyMUMPS(F, b)
print("resid", max(np.linalg.norm(A @ x[:, i] - b[:, i]) / np.linalg.norm(b[:, i]) for i in range(nrhs)))
