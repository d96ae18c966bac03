"""HMC driver — restates `HMCMT/src/HMCSampler/HMCSampler.jl`, `HMCStruct/HMCStruct.jl` and
`HMCUtility/HMCUtility.jl` (test infrastructure).

Random draws are *injected* (`RandomStreams`) because Julia's `rand`/`randn` streams cannot be
reproduced outside Julia; the draw order per sample is the reference's:
`rand(Lmin:Lmax)`, `rand()`, `randn(n)`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np
import scipy.sparse as sp

from . import operators as ops
from .fileio import (TensorMesh2D, MTData, HMCPrior, setupTensorMesh2D, readEMModel2D,
                     readMT2DData, parse_startup)
from .forward import MT2DFwdSolver, default_factor
from .sensitivity import compJacTMatVec


@dataclass
class InvDataModel:
    """`InvDataModel` HMCStruct.jl:75-91."""
    obsData: np.ndarray
    dataW: np.ndarray            # diagonal of Wd
    strModel: np.ndarray
    refModel: np.ndarray
    activeCell: sp.csr_matrix
    bgModel: np.ndarray
    Wm: sp.csr_matrix


def setActiveElement(sigma, sigFix, fixIndex=None):
    """`setActiveElement` HMCUtility.jl:217-258 (exact float compare against sigFix)."""
    sigma = np.asarray(sigma)
    n = len(sigma)
    inaInd = np.zeros(n, dtype=np.int64)
    bg = np.zeros(n)
    for s in sigFix:
        ind = sigma == s
        if not ind.any():
            continue
        inaInd += ind
        bg[ind] += s
    if fixIndex is not None and len(fixIndex):
        inaInd[fixIndex] = 1
        bg[fixIndex] = sigma[fixIndex]
    aInd = np.nonzero(inaInd == 0)[0]
    activeCell = sp.identity(n, format="csc", dtype=np.float64)[:, aInd].tocsr()
    return activeCell, bg


def compDataWeightMat(obsData, dataError):
    """`compDataWeightMat` HMCUtility.jl:168-190 -> diagonal of Wd."""
    return 1.0 / np.abs(np.asarray(dataError, dtype=np.float64))


def getDataMisfit(dataRes):
    """`getDataMisfit` HMCUtility.jl:200-207."""
    return float(np.real(0.5 * np.vdot(dataRes, dataRes)))


def setupInverseDataModel(mesh: TensorMesh2D, sigFix, obsData, dataErr, fixIndex=None) -> InvDataModel:
    """`setupInverseDataModel` HMCStruct.jl:99-125."""
    activeCell, bg = setActiveElement(mesh.sigma, sigFix, fixIndex)
    dataW = compDataWeightMat(obsData, dataErr)
    strModel = np.log(activeCell.T @ np.asarray(mesh.sigma))
    cGrad = ops.getCellGradient2D(mesh.yLen, mesh.zLen) @ activeCell
    Wm = (cGrad.T @ cGrad).tocsr()
    return InvDataModel(np.asarray(obsData), dataW, strModel, strModel.copy(), activeCell, bg, Wm)


def readstartupFile(startupfile: str, base_dir: str = "."):
    """`readstartupFile` readstartupFile.jl:4-103."""
    import os
    datafile, modelfile, sigmin, sigmax, sigfix, prior = parse_startup(startupfile)
    data, obs, err = readMT2DData(os.path.join(base_dir, datafile))
    mesh = readEMModel2D(os.path.join(base_dir, modelfile))
    setupTensorMesh2D(mesh)
    inv = setupInverseDataModel(mesh, sigfix, obs, err)
    return mesh, data, inv, prior


def compDataGradient(mesh, data, inv: InvDataModel, prior: HMCPrior, factor_fn=default_factor, dense_bc=False):
    """4-argument `compDataGradient` HMCSampler.jl:277-330 -> (predData, dataMisfit, dataGrad)."""
    sig_act = np.exp(inv.strModel)                      # modelTransform HMCUtility.jl:69-77
    mesh.sigma = inv.activeCell @ sig_act + inv.bgModel
    pred, fwd = MT2DFwdSolver(mesh, data, factor_fn)
    res = inv.dataW * (pred - inv.obsData)
    misfit = getDataMisfit(res)
    v = inv.dataW * res
    g = compJacTMatVec(fwd.exTE, fwd.hxTM, v, mesh, data, inv.activeCell, fwd.AinvTE, fwd.AinvTM,
                       dense_bc=dense_bc)
    return pred, misfit, sig_act * g


def setMassMatrix(inv, prior):
    """`setMassMatrix` HMCSampler.jl:463-489 -> (invM, sqrtM): None, None for massType "diagonal" (identity, scaling = 1.0,
    :81-83); else the dense Cholesky of Wm: sqrtM = L, invM = inv(L)' * inv(L)."""
    if prior.massType == "diagonal":
        return None, None
    L = np.linalg.cholesky(np.asarray(inv.Wm.todense(), dtype=np.float64))
    Linv = np.linalg.inv(L)
    return Linv.T @ Linv, L


def getKineticEnergy(p, invM=None):
    """`getKineticEnergy` HMCSampler.jl:407-415."""
    return 0.5 * float(np.dot(p, p if invM is None else invM @ p))


def getHamiltonian(data, mesh, inv, prior, momentum, factor_fn=default_factor, invM=None):
    """`getHamiltonian` HMCSampler.jl:358-397 -> (dataMisfit, kp, hmp, mnorm, predData).
    Uses mesh.sigma as left by the last compDataGradient / updateStartModel."""
    pred, _ = MT2DFwdSolver(mesh, data, factor_fn)
    dm = getDataMisfit(inv.dataW * (pred - inv.obsData))
    kp = getKineticEnergy(momentum, invM)
    mprior = inv.strModel - inv.refModel
    mnorm = 0.5 * float(mprior @ (inv.Wm @ mprior)) * prior.regParam
    return dm, kp, dm + kp + mnorm, mnorm, pred


def checkParameterBound(model, momentum, prior):
    """`checkParameterBound!` HMCSampler.jl:515-559 (reflection at ln sigma bounds)."""
    lo, hi = np.log(prior.sigBounds[0]), np.log(prior.sigBounds[1])
    for k in range(len(model)):
        if lo <= model[k] <= hi:
            continue
        niter = 0
        while True:
            niter += 1
            if model[k] < lo:
                model[k] = 2.0 * lo - model[k]
                momentum[k] *= -1.0
            if model[k] > hi:
                model[k] = 2.0 * hi - model[k]
                momentum[k] *= -1.0
            if lo <= model[k] <= hi:
                break
            if niter >= 500:      # the reference only prints and loops forever; bail out here
                raise RuntimeError(f"constraints for {k} is not fulfilled")
    return model, momentum


def clip_momentum(z):
    """`getMomentumVector` HMCSampler.jl:441-453 applied to an injected randn draw."""
    z = np.array(z, dtype=np.float64)
    return np.clip(z, -2.5, 2.5)


def proposeLeapfrog(model, momentum, mesh, data, inv, prior, intstep, factor_fn=default_factor, trace=None, invM=None):
    """`proposeLeapfrog` HMCSampler.jl:206-269 with `intstep` injected (rand(t1:t2), :233); invM: getKineticGradient (:424-431)."""
    inv.strModel = model.copy()
    _, _, g = compDataGradient(mesh, data, inv, prior, factor_fn)
    g = g + (inv.Wm @ (model - inv.refModel)) * prior.regParam
    dt = prior.dt
    p = momentum - 0.5 * dt * g
    m = model.copy()
    for k in range(1, intstep + 1):
        dm = dt * (p if invM is None else invM @ p)
        dmMax = np.max(np.abs(dm))
        if dmMax > 3.0:
            dm = dm / dmMax * 3.0
        m = m + dm
        m, p = checkParameterBound(m, p, prior)
        inv.strModel = m.copy()
        pred, mis, g = compDataGradient(mesh, data, inv, prior, factor_fn)
        g = g + (inv.Wm @ (m - inv.refModel)) * prior.regParam
        if trace is not None:
            trace.append(dict(m=m.copy(), misfit=mis, grad=g.copy()))
        delta = dt * g
        p = p - delta if k < intstep else p - 0.5 * delta
    return m, p


@dataclass
class RandomStreams:
    """Injected draws: start-model uniform, initial momentum normals, and per sample
    (leapfrog step count, accept uniform, momentum normals)."""
    u_start: float
    z_init: np.ndarray
    intsteps: np.ndarray
    u_accept: np.ndarray
    z_momentum: np.ndarray      # (nsamples, nparam)


def make_streams(seed, nparam, nsamples, timestep):
    rng = np.random.default_rng(seed)
    return RandomStreams(float(rng.random()), rng.standard_normal(nparam),
                         rng.integers(timestep[0], timestep[1] + 1, size=nsamples),
                         rng.random(nsamples), rng.standard_normal((nsamples, nparam)))


def runHMCSampler(mesh, data, inv: InvDataModel, prior: HMCPrior, streams: RandomStreams,
                  nsamples=None, factor_fn=default_factor):
    """`runHMCSampler` HMCSampler.jl:72-196 -> (hmcmodel[nparam,nsamples], stats dict, hmcdata)."""
    nparam, ndata = len(inv.strModel), len(inv.obsData)
    nsamples = prior.totalsamples if nsamples is None else nsamples
    invM, sqrtM = setMassMatrix(inv, prior)                 # :81-86
    draw = (lambda z: clip_momentum(z)) if sqrtM is None else (lambda z: sqrtM @ clip_momentum(z))      # getMomentumVector :441-453
    cur_m = np.array(inv.strModel, dtype=float).copy()      # hmcParamCurrent.rhomodel = copy(invParam.strModel) (:87): the
    #                                                         model-file model, taken BEFORE strModel is replaced below (:100-109)
    cur_p = draw(streams.z_init)
    sigma0 = inv.strModel[0]          # unique(strModel)[1]: Julia's unique keeps first-appearance order (:100-101)
    rho0 = 1.0 / np.exp(sigma0)
    rhoref = np.round(rho0 * 0.5 + (rho0 * 1.5 - rho0 * 0.5) * streams.u_start)
    strModel = np.log(np.ones(nparam) / rhoref)
    inv.strModel = strModel.copy()
    inv.refModel = strModel.copy()
    mesh.sigma = inv.activeCell @ np.exp(inv.strModel) + inv.bgModel       # updateStartModel :834-849
    startD, startK, startH, startM, pred = getHamiltonian(data, mesh, inv, prior, cur_p, factor_fn, invM)
    hmcmodel = np.zeros((nparam, nsamples))
    hmcdata = np.zeros((ndata, nsamples + 1), dtype=np.complex128)
    hmstats = np.zeros((4, nsamples + 1))
    accept = np.zeros(nsamples, dtype=bool)
    hmstats[:, 0] = [startD, startM, startK, startH]
    hmcdata[:, 0] = pred
    nAccept = nReject = 0
    for it in range(1, nsamples + 1):
        pm, pp = proposeLeapfrog(cur_m, cur_p, mesh, data, inv, prior, int(streams.intsteps[it - 1]), factor_fn, invM=invM)
        finD, finK, finH, finM, pred = getHamiltonian(data, mesh, inv, prior, pp, factor_fn, invM)
        hdif = startH - finH
        if hdif > 0 or streams.u_accept[it - 1] < np.exp(hdif):
            cur_m, cur_p = pm.copy(), pp.copy()
            startD, startM = finD, finM
            nAccept += 1
            accept[it - 1] = True
            hmcdata[:, it] = pred
        else:
            nReject += 1
            hmcdata[:, it] = hmcdata[:, it - 1]
        cur_p = draw(streams.z_momentum[it - 1])
        startK = getKineticEnergy(cur_p, invM)
        startH = startD + startM + startK
        hmstats[:, it] = [startD, startM, startK, startH]
        hmcmodel[:, it - 1] = cur_m
    stats = dict(nAccept=nAccept, nReject=nReject, acceptstats=accept, hmstats=hmstats)
    return hmcmodel, stats, hmcdata
