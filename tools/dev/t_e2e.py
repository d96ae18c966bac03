import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from hmcmt2d_b200 import api, synthetic
mesh, data, inv, prior = synthetic.make_problem(200, 100, 30)
pl = api.Plan(mesh, data, inv, prior)
m = synthetic.stress_model(inv)
for _ in range(5): pl.forward_gradient(m)
t0 = time.perf_counter()
for _ in range(50): pl.forward_gradient(m)
t1 = time.perf_counter()
print("forward_gradient per call %.3f ms" % ((t1 - t0) / 50 * 1e3))
p = np.zeros_like(m)
pl.set_state(m, p, m)
pl.leapfrog_steps_device(prior.dt, 5); pl.sync()
pl.timer_start(); pl.leapfrog_steps_device(prior.dt, 50); ms = pl.timer_stop()
print("device step %.3f ms" % (ms / 50))
import ctypes as C
from hmcmt2d_b200 import lib as _lib
mm = pl._m(m); pred = np.zeros((1, pl.nData), dtype=np.complex128); phi = np.zeros(1); g = np.zeros((1, pl.nAC))
a = (_lib.f64(mm), pred.ctypes.data_as(C.POINTER(C.c_double)), _lib.f64(phi), _lib.f64(g))
t0 = time.perf_counter()
for _ in range(50): pl.L.hmcmt_forward_gradient(pl.h, *a)
t1 = time.perf_counter()
print("raw C call per call %.3f ms" % ((t1 - t0) / 50 * 1e3))
