"""Dev: large-bandwidth path through the MUMPS-shim ABI (random stencil and full-band matrices) and through the plan."""
import sys, time, numpy as np, scipy.sparse as sp
sys.path.insert(0, '.')
from hmcmt2d_b200 import lib
rng = np.random.default_rng(2)
def relres(A, x, b): return np.linalg.norm(A @ x - b) / np.linalg.norm(b)
for nl, nf in [(4, 105), (7, 130), (5, 200), (4, 299), (3, 320)]:
    N = nl * nf
    d = 4 + rng.random(N) + 1j * rng.random(N)
    e1, e2 = -rng.random(N), -rng.random(N)
    e1[np.arange(N) % nf == 0] = 0
    A = sp.diags([d, e1[1:], e1[1:], e2[nf:], e2[nf:]], [0, -1, 1, -nf, nf], format="csc")
    rhs = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    t0 = time.time(); x = lib.solveMUMPS(A, rhs, 1)
    print('stencil', nl, nf, 'relres', relres(A, x, rhs), 'time', round(time.time() - t0, 3), flush=True)
# full band, b = 120, N not a multiple of 32
N, b = 1003, 120
M = sp.random(N, N, density=0.0, format='lil')
offs = list(range(1, b + 1))
diags = [8 * b * (1 + rng.random(N)) + 1j * rng.random(N)] + [rng.standard_normal(N - o) + 1j * rng.standard_normal(N - o) for o in offs]
A = sp.diags([diags[0]] + diags[1:] + diags[1:], [0] + offs + [-o for o in offs], format='csc')
rhs = rng.standard_normal((N, 3)) + 1j * rng.standard_normal((N, 3))
x = lib.solveMUMPS(A, rhs, 2)
print('full band', relres(A, x[:, 0], rhs[:, 0]), relres(A, x[:, 2], rhs[:, 2]), flush=True)
if len(sys.argv) > 1:
    from hmcmt2d_b200 import api, synthetic
    ny, nz, nf = (int(a) for a in sys.argv[1:4])
    mesh, data, inv, prior = synthetic.make_problem(ny, nz, nf, nRx=10)
    m = synthetic.stress_model(inv)
    pl = api.Plan(mesh, data, inv, prior)
    print('N', pl.info(0), 'b', pl.info(4), 'T', pl.info(5), 'S', pl.info(6), flush=True)
    t0 = time.time(); pred, phi, g = pl.forward_gradient(m); t1 = time.time()
    pred, phi, g = pl.forward_gradient(m); t2 = time.time()
    print('grad ok', phi, np.abs(g).max(), 'first', round(t1 - t0, 3), 'second', round(t2 - t1, 3), flush=True)
    np.savez('gpurun_out/big_%d_%d_%d.npz' % (ny, nz, nf), pred=pred, phi=phi, g=g)
