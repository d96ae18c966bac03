import sys, ctypes as C, numpy as np, scipy.sparse as sp
sys.path.insert(0, '/root/repo')
from hmcmt2d_b200 import lib
L = lib.load()
rng = np.random.default_rng(0)
nl, nf = 4, 10
N = nl*nf
d = 4 + rng.random(N) + 1j*rng.random(N)
e1 = -rng.random(N); e2 = -rng.random(N)
e1[np.arange(N) % nf == 0] = 0
A = sp.diags([d, e1[1:], e1[1:], e2[nf:], e2[nf:]], [0, -1, 1, -nf, nf], format='csc')
f = lib.factorMUMPS(A, 1)
dims = np.zeros(4, dtype=np.int64)
L.hmcmt_debug_get_factor.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
L.hmcmt_debug_get_factor(f.ptr, None, None, dims.ctypes.data)
n, b, T, S = dims
R = 8*T
pan = np.zeros((S, 2, 2, R, 4)); ainv = np.zeros((S, 64), dtype=complex)
L.hmcmt_debug_get_factor(f.ptr, pan.ctypes.data, ainv.ctypes.data, dims.ctypes.data)
print('dims', dims)
# emulate
Ad = A.toarray()
def entry(g, c):
    if g >= N or c >= N: return 1.0 if g == c else 0.0
    return Ad[g, c]
W = np.zeros((R, R), dtype=complex); glob = np.arange(T)
def fill(X):
    for Y in range(T):
        for rr in range(8):
            for cc in range(8):
                v = entry(glob[X]*8+rr, glob[Y]*8+cc); W[X*8+rr, Y*8+cc] = v; W[Y*8+cc, X*8+rr] = v
for X in range(T): fill(X)
for s in range(S):
    p = s % T; ps = slice(p*8, p*8+8)
    raw = W[:, ps].copy(); A11 = raw[ps, :].copy()
    Ainv = np.linalg.inv(A11)
    # device raw image
    draw = np.zeros((R, 8), dtype=complex)
    for c in range(8):
        draw[:, c] = pan[s, 0, c >> 2, :, c & 3] + 1j*pan[s, 1, c >> 2, :, c & 3]
    e_raw = np.abs(draw - raw).max(); e_inv = np.abs(ainv[s].reshape(8, 8) - Ainv).max()
    print('step', s, 'raw err', e_raw, 'ainv err', e_inv)
    if e_raw > 1e-10:
        bad = np.argwhere(np.abs(draw - raw) > 1e-10)
        print('  bad rows/cols (first 12):', bad[:12].tolist())
        print('  dev', draw[bad[0][0], bad[0][1]], 'ref', raw[bad[0][0], bad[0][1]])
    raw2 = raw.copy(); raw2[ps, :] = 0
    W -= raw2 @ Ainv @ raw2.T
    glob[p] = s + T; W[ps, :] = 0; W[:, ps] = 0; fill(p)
