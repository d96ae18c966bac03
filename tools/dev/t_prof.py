import sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo')
from hmcmt2d_b200 import api, synthetic
mesh, data, inv, prior = synthetic.make_problem(200, 100, 30)
pl = api.Plan(mesh, data, inv, prior)
m = synthetic.stress_model(inv)
pl.forward_gradient(m); pl.forward_gradient(m)
out = np.zeros(32, dtype=np.uint64)
pl.L.hmcmt_debug_prof(out.ctypes.data_as(C.c_void_p))
S = pl.info(6) // 2   # macro-steps of one half (FM_OWN launch, CTA 0)
print('per-step cycles, tile warp 0:', dict(zip(['waitINV','Mprime+y','BAR_M','recycle','pass1+dump+arrive','pass2'], (out[0:6] / S).round(0))), 'sum', out[0:6].sum() / S)
print('per-step cycles, tile warp 7:', dict(zip(['waitINV','Mprime+y','BAR_M','recycle','pass1+dump+arrive','pass2'], (out[8:14] / S).round(0))), 'sum', out[8:14].sum() / S)
print('per-step cycles, factor warp:', dict(zip(['(waitRAW)','z+arriveINV','ring+early-update','GJ','publish+store','ringwrite','bar'], (out[16:23] / S).round(0))), 'sum', out[16:23].sum() / S)
