import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from hmcmt2d_b200 import api, synthetic
ny, nz, nf = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (200, 100, 30)
mesh, data, inv, prior = synthetic.make_problem(ny, nz, nf)
m = synthetic.stress_model(inv)
pl = api.Plan(mesh, data, inv, prior)
print('N', pl.info(0), 'b', pl.info(4), 'T', pl.info(5), 'S', pl.info(6), 'nsys', pl.info(7), 'factor MB/system', pl.info(9) / 1e6)
t = time.time(); pred, phi, g = pl.forward_gradient(m); print('first', time.time() - t, phi)
for _ in range(3):
    t = time.time(); pred, phi, g = pl.forward_gradient(m); print('e2e step', time.time() - t, phi, np.abs(g).max())
pl.set_state(m, np.zeros_like(m), m)
pl.kernel_time(True)
pl.leapfrog_steps_device(0.0, 2); pl.sync()
pl.kernel_time(True)
n = 10
pl.timer_start(); pl.leapfrog_steps_device(0.0, n); ms = pl.timer_stop()
fms, fl = pl.kernel_time(True)
print('device step ms', ms / n, 'factor kernel ms', fms / fl, 'steps/s', 1000 * n / ms)
N, b = pl.info(0), pl.info(4)
flops = 4.0 * N * b * b * pl.info(7)
print('factor algorithmic TFLOP/s', flops / (fms / fl * 1e-3) / 1e12)
