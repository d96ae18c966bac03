import sys, numpy as np
sys.path.insert(0, '/root/repo')
from hmcmt2d_b200 import api
from oracle import sampler as osamp, forward as ofwd, sensitivity as osens, operators as oops
ex = '/root/repo/tests/golden/dprism3d'
mesh, data, inv, prior = api.readstartupFile(ex + '/startupfile')
omesh, odata, oinv, oprior = osamp.readstartupFile(ex + '/startupfile', ex)
rng = np.random.default_rng(1)
m = np.log(0.01) + 0.7 * rng.standard_normal(len(inv.strModel))
inv.strModel = m.copy(); oinv.strModel = m.copy()
pred, phi, g = api.compDataGradient(mesh, data, inv, prior)
pl = api._plan_for(mesh, data, inv, prior)
omesh.sigma = oinv.activeCell @ np.exp(m) + oinv.bgModel
ny, nz = omesh.gridSize
ii, io = oops.getBoundaryIndex(ny, nz)
for mode, nm in [(0, 'TE'), (1, 'TM')]:
    coe = ofwd.assemble_mode(omesh, mode == 0, ii, io)
    for f in [0, 5, 10]:
        colptr, rowval, nzval, rhs, bc = pl.export_system(mode, f)
        om = 2 * np.pi * data.freqs[f]
        Aii = (coe.rAii + 1j * om * coe.iAii).tocsc(); Aii.sort_indices()
        obc = ofwd.getBoundaryMT2DTE(data.freqs[f], omesh.yLen, omesh.zLen, omesh.sigma) if mode == 0 else ofwd.getBoundaryMT2DTM(data.freqs[f], omesh.yLen, omesh.zLen, omesh.sigma)
        Aio = (coe.rAio + 1j * om * coe.iAio)
        orhs = -(Aio @ obc)
        print(nm, f, 'pattern', np.array_equal(colptr, Aii.indptr + 1), np.array_equal(rowval, Aii.indices + 1),
              'vals', np.abs(nzval - Aii.data).max() / np.abs(Aii.data).max(),
              'bc', np.abs(bc - obc).max(), 'at', np.argmax(np.abs(bc - obc)), 'rhs', np.abs(rhs - orhs).max() / np.abs(orhs).max())
        d = np.abs(bc - obc)
        sec = {'top': slice(0, ny + 1), 'left': slice(ny + 1, ny + nz + 1), 'right': slice(ny + nz + 1, ny + 2 * nz + 1), 'bot': slice(ny + 2 * nz + 1, None)}
        print('    ', {k: float(d[v].max()) for k, v in sec.items()}, 'max|bc bot|', np.abs(obc[sec['bot']]).max())
