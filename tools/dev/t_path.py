import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from hmcmt2d_b200 import api
from oracle import sampler as osamp, forward as ofwd, sensitivity as osens
ex = sys.argv[1] if len(sys.argv) > 1 else '/root/repo/tests/golden/dprism3d'
mesh, data, inv, prior = api.readstartupFile(ex + '/startupfile')
omesh, odata, oinv, oprior = osamp.readstartupFile(ex + '/startupfile', ex)
rng = np.random.default_rng(1)
m = np.log(0.01) + 0.7 * rng.standard_normal(len(inv.strModel))
inv.strModel = m.copy(); oinv.strModel = m.copy()
t = time.time(); pred, phi, g = api.compDataGradient(mesh, data, inv, prior); print('gpu first call', time.time() - t)
t = time.time(); pred, phi, g = api.compDataGradient(mesh, data, inv, prior); print('gpu second call', time.time() - t)
t = time.time(); parts = {}
opred, ophi, og = osamp.compDataGradient(omesh, odata, oinv, oprior); print('oracle', time.time() - t)
print('pred rel err', np.abs(pred - opred).max() / np.abs(opred).max(), 'max elementwise rel', (np.abs(pred - opred) / np.abs(opred)).max())
print('phi', phi, ophi, abs(phi - ophi) / abs(ophi))
print('grad rel err (max|g|)', np.abs(g - og).max() / np.abs(og).max())
bad = np.argsort(-np.abs(g - og))[:5]; print('worst', bad, g[bad], og[bad])
# fields
p2, fwd = api.MT2DFwdSolver(mesh.__class__(mesh.yLen, mesh.zLen, mesh.airLayer, mesh.gridSize, mesh.origin, omesh.sigma), data)
op2, ofw = ofwd.MT2DFwdSolver(omesh, odata)
print('MT2DFwdSolver pred', np.abs(p2 - op2).max() / np.abs(op2).max())
print('exTE', np.abs(fwd.exTE - ofw.exTE).max() / np.abs(ofw.exTE).max(), 'hxTM', np.abs(fwd.hxTM - ofw.hxTM).max() / np.abs(ofw.hxTM).max())
v = rng.standard_normal(len(op2)) + 1j * rng.standard_normal(len(op2))
gs = api.compJacTMatVec(fwd.exTE, fwd.hxTM, v, mesh, data, None, fwd.AinvTE, fwd.AinvTM)
ogs = osens.compJacTMatVec(ofw.exTE, ofw.hxTM, v, omesh, odata, oinv.activeCell, ofw.AinvTE, ofw.AinvTM)
print('jtvec', np.abs(gs - ogs).max() / np.abs(ogs).max())
