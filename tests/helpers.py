"""Shared builders for the tests (oracle-side objects; the product side reads the same files / arrays)."""
import os

import numpy as np

from oracle import fileio as ofio
from oracle import sampler as osamp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def tiny_problem(seed=3, nF=3, ny_core=4):
    """8 x 8 cell mesh (3 air layers), 3 frequencies, 4 receivers, random log-normal earth."""
    rng = np.random.default_rng(seed)
    air = np.array([100.0, 1000.0, 10000.0])
    ylen = np.concatenate([[3200.0, 800.0], np.full(ny_core, 200.0), [800.0, 3200.0]])
    zlen = np.concatenate([air[::-1], [100.0, 100.0, 200.0, 400.0, 1600.0]])
    ny, nz = len(ylen), len(zlen)
    sig = np.concatenate([np.full(ny * 3, 1e-8), np.exp(np.log(0.01) + 0.7 * rng.standard_normal(ny * (nz - 3)))])
    mesh = ofio.TensorMesh2D(ylen, zlen, air, (ny, nz), np.array([4000.0 + 200.0, air.sum()]), sig)
    freqs = np.array([10.0, 1.0, 0.1])[:nF]
    nRx = 4
    rx = np.stack([np.array([-150.0, 30.0, 100.0, 260.0]), np.zeros(nRx)], 1)
    f, r, c = np.meshgrid(np.arange(1, nF + 1), np.arange(1, nRx + 1), np.arange(1, 3), indexing="ij")
    data = ofio.MTData(rx, freqs, "Impedance", ["ZXY", "ZYX"], r.ravel().astype(np.int64), f.ravel().astype(np.int64),
                       c.ravel().astype(np.int64), np.ones(nF * nRx * 2, bool), True, True)
    nd = nF * nRx * 2
    obs = (rng.standard_normal(nd) + 1j * rng.standard_normal(nd)) * 1e-2
    err = 0.05 * np.abs(obs) + 1e-4
    inv = osamp.setupInverseDataModel(mesh, [1e-8], obs, err)
    prior = ofio.HMCPrior(sigBounds=[1e-4, 1.0], dt=0.03, timestep=[2, 3], regParam=1.0)
    return mesh, data, inv, prior


def load_example(name):
    d = os.path.join(GOLDEN, name)
    return osamp.readstartupFile(os.path.join(d, "startupfile"), d)


def to_product(mesh, data, inv, prior):
    """Build the product-side (hmcmt2d_b200) objects from oracle-side ones (same arrays)."""
    from hmcmt2d_b200 import api, fileio
    pm = fileio.TensorMesh2D(np.array(mesh.yLen), np.array(mesh.zLen), np.array(mesh.airLayer), tuple(mesh.gridSize),
                             np.array(mesh.origin), np.array(mesh.sigma))
    pd = fileio.MTData(np.array(data.rxLoc), np.array(data.freqs), data.dataType, list(data.dataComp), np.array(data.rxID),
                       np.array(data.freqID), np.array(data.dtID), np.array(data.dataID), data.compTE, data.compTM)
    pi = api.InvDataModel(np.array(inv.obsData), np.array(inv.dataW), np.array(inv.strModel), np.array(inv.refModel),
                          inv.activeCell.copy(), np.array(inv.bgModel), inv.Wm.copy(), 1.0 / np.array(inv.dataW))
    pp = fileio.HMCPrior(burninsamples=prior.burninsamples, totalsamples=prior.totalsamples, sigBounds=list(prior.sigBounds),
                         sigmastd=prior.sigmastd, dt=prior.dt, timestep=list(prior.timestep), regParam=prior.regParam)
    return pm, pd, pi, pp


def to_oracle(mesh, data, inv, prior):
    """Oracle-side objects from product-side ones (hmcmt2d_b200.synthetic / fileio), same arrays."""
    omesh = ofio.TensorMesh2D(np.array(mesh.yLen), np.array(mesh.zLen), np.array(mesh.airLayer), tuple(mesh.gridSize),
                              np.array(mesh.origin), np.array(mesh.sigma))
    od = ofio.MTData(np.array(data.rxLoc), np.array(data.freqs), data.dataType, list(data.dataComp), np.array(data.rxID),
                     np.array(data.freqID), np.array(data.dtID), np.array(data.dataID), data.compTE, data.compTM)
    oinv = osamp.setupInverseDataModel(omesh, [1e-8], np.array(inv.obsData), np.array(inv.dataErr))
    oinv.strModel = np.array(inv.strModel)
    oprior = ofio.HMCPrior(burninsamples=prior.burninsamples, totalsamples=prior.totalsamples, sigBounds=list(prior.sigBounds),
                           sigmastd=prior.sigmastd, dt=prior.dt, timestep=list(prior.timestep), regParam=prior.regParam)
    return omesh, od, oinv, oprior
