import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from hmcmt2d_b200 import api, synthetic
ny, nz, nf = (int(a) for a in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else 'fwd'
mesh, data, inv, prior = synthetic.make_problem(ny, nz, nf, nRx=10)
m = synthetic.stress_model(inv)
pl = api.Plan(mesh, data, inv, prior)
print('N', pl.info(0), 'b', pl.info(4), 'T', pl.info(5), 'S', pl.info(6), flush=True)
if mode == 'fwd':
    pred, ex, hx = pl.forward(m=m)
    print('fwd ok', np.abs(pred).max(), flush=True)
else:
    pred, phi, g = pl.forward_gradient(m)
    print('grad ok', phi, np.abs(g).max(), flush=True)
np.save('/tmp/pred_%s.npy' % mode, pred)
