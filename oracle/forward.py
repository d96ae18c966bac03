"""2-D MT forward problem — restates `HMCMT/src/MTFwdSolver/{MT2DFwdSolver,mt2DTE,mt2DTM,
mt1DField}.jl` literally with scipy.sparse (test infrastructure).
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import operators as ops
from .fileio import TensorMesh2D, MTData, setupTensorMesh2D

MU0 = 4 * np.pi * 1e-7
EPS0 = 8.85 * 1e-12


@dataclass
class CoeffMat:
    """`CoeffMat` MT2DFwdSolver.jl:19-29."""
    rAii: sp.csc_matrix
    iAii: sp.csc_matrix
    rAio: sp.csc_matrix
    iAio: sp.csc_matrix


@dataclass
class MT2DFwdData:
    """`MT2DFwdData` MT2DFwdSolver.jl:44-53 (fields nNode x nFreq, factor handles)."""
    exTE: np.ndarray
    hxTM: np.ndarray
    AinvTE: list
    AinvTM: list
    linearSolver: str = ""


def mt1DAnalyticField(freq, sigma, zNode, compH=False):
    """`mt1DAnalyticField` mt1DField.jl:23-98.  Layered half-space, e^{+iwt}."""
    sigma = np.asarray(sigma, dtype=np.float64)
    zNode = np.asarray(zNode, dtype=np.float64)
    assert len(sigma) == len(zNode) - 1
    eTop = 1.0 + 0j
    omega = 2 * np.pi * freq
    omu0 = omega * MU0
    sigma = np.concatenate([sigma, sigma[-1:]])
    nLayer = len(zNode)
    zLen = np.diff(zNode)

    k = np.sqrt(MU0 * EPS0 * omega ** 2 - MU0 * sigma[-1] * omega * 1j + 0j)
    ztmp = omega * MU0 / k
    for j in range(nLayer - 2, -1, -1):
        k = np.sqrt(MU0 * EPS0 * omega ** 2 - MU0 * sigma[j] * omega * 1j + 0j)
        zp = omega * MU0 / k
        th = np.tanh(k * zLen[j] * 1j)
        ztmp = zp * (ztmp + zp * th) / (zp + ztmp * th)
    z0 = ztmp

    eLayer = np.zeros((2, nLayer), dtype=np.complex128)
    # NB `k` here is the wavenumber of the top layer (last loop iteration), mt1DField.jl:62-63
    eLayer[0, 0] = 0.5 * eTop * (1 - omega * MU0 / (z0 * k))
    eLayer[1, 0] = 0.5 * eTop * (1 + omega * MU0 / (z0 * k))

    ka = np.sqrt(MU0 * EPS0 * omega ** 2 - MU0 * sigma * omega * 1j + 0j)
    with np.errstate(over="ignore", invalid="ignore"):
        for i in range(nLayer - 1):
            kr = ka[i] / ka[i + 1]
            eu = np.exp(ka[i] * zLen[i] * 1j) * eLayer[0, i]
            ed = np.exp(-ka[i] * zLen[i] * 1j) * eLayer[1, i]
            eLayer[0, i + 1] = 0.5 * ((1 + kr) * eu + (1 - kr) * ed)
            eLayer[1, i + 1] = 0.5 * ((1 - kr) * eu + (1 + kr) * ed)
            e2 = abs(eLayer[0, i + 1] + eLayer[1, i + 1])
            e1 = abs(eLayer[0, i] + eLayer[1, i])
            if e2 - e1 > 0 or np.isnan(e2):
                eLayer[:, i + 1:] = 0.0
                break
    eField = eLayer.sum(axis=0)
    if compH:
        hField = (-ka * eLayer[0] + ka * eLayer[1]) / omu0
        return eField, hField
    return eField


def mt1DImpedance(freqs, sigma, zNode):
    """Surface impedance of a layered model, `mt1DImpedance` mt1DField.jl:115-165 (Z only).

    sigma has one entry per zNode (the last one is the half-space).
    """
    sigma = np.asarray(sigma, dtype=np.float64)
    zNode = np.asarray(zNode, dtype=np.float64)
    h = np.diff(zNode)
    out = np.zeros(len(freqs), dtype=np.complex128)
    for i, f in enumerate(freqs):
        omega = 2 * np.pi * f
        k = np.sqrt(MU0 * EPS0 * omega ** 2 - MU0 * sigma[-1] * omega * 1j + 0j)
        z = omega * MU0 / k
        for j in range(len(sigma) - 2, -1, -1):
            k = np.sqrt(MU0 * EPS0 * omega ** 2 - MU0 * sigma[j] * omega * 1j + 0j)
            zp = omega * MU0 / k
            th = np.tanh(k * h[j] * 1j)
            z = zp * (z + zp * th) / (zp + z * th)
        out[i] = z
    return out


def _boundary(freq, yLen, zLen, sigma, useH):
    """Common body of `getBoundaryMT2DTE` mt2DTE.jl:100-134 / `getBoundaryMT2DTM` mt2DTM.jl:100-134."""
    ny, nz = len(yLen), len(zLen)
    zNode = np.concatenate([[0.0], np.cumsum(zLen)])
    sigma2D = np.asarray(sigma).reshape(nz, ny)          # sigma2D[k, j]
    nb = 2 * (ny + nz)
    bc = np.zeros(nb, dtype=np.complex128)
    bc[0:ny + 1] = 1.0

    def prof(s1d):
        if useH:
            _, f = mt1DAnalyticField(freq, s1d, zNode, True)
        else:
            f = mt1DAnalyticField(freq, s1d, zNode)
        return f

    fb = prof(sigma2D[:, 0])
    fb = fb / fb[0]
    bc[ny + 1:ny + nz + 1] = fb[1:]
    fb = prof(sigma2D[:, -1])
    fb = fb / fb[0]
    bc[ny + nz + 1:ny + 2 * nz + 1] = fb[1:]
    for i in range(1, ny):      # reference i = 2..ny  ->  bc[ny+2nz+i] (1-based)
        s1d = (sigma2D[:, i - 1] * yLen[i - 1] + sigma2D[:, i] * yLen[i]) / (yLen[i - 1] + yLen[i])
        fb = prof(s1d)
        bc[ny + 2 * nz + i] = fb[-1] / fb[0]
    return bc


def getBoundaryMT2DTE(freq, yLen, zLen, sigma):
    return _boundary(freq, yLen, zLen, sigma, False)


def getBoundaryMT2DTM(freq, yLen, zLen, sigma):
    return _boundary(freq, yLen, zLen, sigma, True)


def _interp_rx(rxLoc, yNode, f0, g0):
    """Un-normalised linear interpolation, mt2DTE.jl:196-207 / mt2DTM.jl:196-207."""
    nRx = rxLoc.shape[0]
    fr = np.zeros(nRx, dtype=np.complex128)
    gr = np.zeros(nRx, dtype=np.complex128)
    for ir in range(nRx):
        rxY = rxLoc[ir, 0]
        idx = np.nonzero(yNode > rxY)[0]
        if len(idx) == 0:
            raise ValueError("The receiver location seems to be out of range!")
        i = idx[0]
        dy1 = rxY - yNode[i - 1]
        dy2 = yNode[i] - rxY
        fr[ir] = f0[i - 1] * dy2 + f0[i] * dy1
        gr[ir] = g0[i - 1] * dy2 + g0[i] * dy1
    return fr, gr


def compFieldsAtRxTE(omega, rxLoc, yNode, zLen1, sigma1, Er01):
    """`compFieldsAtRxTE` mt2DTE.jl:153-210.  Er01: (ny+1, 2)."""
    yLen = np.diff(yNode)
    ny = len(yLen)
    mu = MU0 * np.ones(ny)
    Ex0 = Er01[:, 0]
    Bz0 = (ops.ddx(ny) @ Er01[:, 0]) / yLen / (1j * omega)
    Bz1 = (ops.ddx(ny) @ Er01[:, 1]) / yLen / (1j * omega)
    HzQ = (0.75 * Bz0 + 0.25 * Bz1) / mu
    HyH = -(Er01[1:-1, 1] - Er01[1:-1, 0]) / zLen1 / (1j * omega * MU0)
    ExQ = 0.75 * Er01[1:-1, 0] + 0.25 * Er01[1:-1, 1]
    avm = ops.av(ny - 1)
    sigma1v = (avm @ (sigma1 * yLen)) / (avm @ yLen)
    dHzQ = (ops.ddx(ny - 1) @ HzQ) / (avm @ yLen)
    Hy0 = np.zeros(ny + 1, dtype=np.complex128)
    Hy0[1:-1] = HyH - (dHzQ - sigma1v * ExQ) * (0.5 * zLen1)
    Hy0[0] = Hy0[1]
    Hy0[-1] = Hy0[-2]
    return _interp_rx(rxLoc, yNode, Ex0, Hy0)       # (Exr, Hyr)


def compFieldsAtRxTM(omega, rxLoc, yNode, zLen1, sigma1, Hr01):
    """`compFieldsAtRxTM` mt2DTM.jl:152-210."""
    yLen = np.diff(yNode)
    ny = len(yLen)
    Hx0 = Hr01[:, 0]
    Jz0 = -(ops.ddx(ny) @ Hr01[:, 0]) / yLen
    Jz1 = -(ops.ddx(ny) @ Hr01[:, 1]) / yLen
    EzQ = (0.75 * Jz0 + 0.25 * Jz1) / sigma1
    JyH = (Hr01[1:-1, 1] - Hr01[1:-1, 0]) / zLen1
    avm = ops.av(ny - 1)
    rho1v = (avm @ ((1.0 / sigma1) * yLen)) / (avm @ yLen)
    EyH = JyH * rho1v
    HxQ = 0.75 * Hr01[1:-1, 0] + 0.25 * Hr01[1:-1, 1]
    dEzQ = (ops.ddx(ny - 1) @ EzQ) / (avm @ yLen)
    Ey0 = np.zeros(ny + 1, dtype=np.complex128)
    Ey0[1:-1] = EyH - (dEzQ + 1j * omega * MU0 * HxQ) * (0.5 * zLen1)
    Ey0[0] = Ey0[1]
    Ey0[-1] = Ey0[-2]
    Eyr, Hxr = _interp_rx(rxLoc, yNode, Ey0, Hx0)
    return Eyr, Hxr


def _resp(omega, num, den, dataType):
    """`compMTRespTE` mt2DTE.jl:240-259 / `compMTRespTM` mt2DTM.jl:224-242."""
    Z = num / den
    if "Impedance" in dataType:
        return np.stack([Z.real, Z.imag], axis=1)
    rho = np.abs(Z) ** 2 / (omega * MU0)
    phs = np.arctan2(Z.imag, Z.real) * 180 / np.pi
    return np.stack([rho, phs], axis=1)


def receiver_row(mesh: TensorMesh2D, rxLoc) -> int:
    """0-based node row of the receivers (mt2DTE.jl:65-67)."""
    zNode = np.concatenate([[0.0], np.cumsum(mesh.zLen)]) - mesh.origin[1]
    hit = np.nonzero(np.abs(zNode - rxLoc[0, 1]) < 0.1)[0]
    return int(hit[0])


def _solve_mode(freq, mesh, coe: CoeffMat, rxLoc, dataType, isTE, factor_fn):
    """`compMT2DTE` mt2DTE.jl:19-83 / `compMT2DTM` mt2DTM.jl:18-83."""
    yLen, zLen, origin, sigma = mesh.yLen, mesh.zLen, mesh.origin, mesh.sigma
    yNode = np.concatenate([[0.0], np.cumsum(yLen)]) - origin[0]
    ny, nz = len(yLen), len(zLen)
    omega = 2 * np.pi * freq
    Aii = (coe.rAii + 1j * omega * coe.iAii).tocsc()
    Aio = (coe.rAio + 1j * omega * coe.iAio).tocsc()
    bc = getBoundaryMT2DTE(freq, yLen, zLen, sigma) if isTE else getBoundaryMT2DTM(freq, yLen, zLen, sigma)
    rhs = -(Aio @ bc)
    Ainv = factor_fn(Aii)
    Fii = Ainv.solve(rhs)
    F2d = np.zeros((nz + 1, ny + 1), dtype=np.complex128)
    F2d[0, :] = bc[0:ny + 1]
    F2d[1:, 0] = bc[ny + 1:ny + nz + 1]
    F2d[1:, -1] = bc[ny + nz + 1:ny + 2 * nz + 1]
    F2d[-1, 1:-1] = bc[ny + 2 * nz + 1:]
    F2d[1:-1, 1:-1] = Fii.reshape(nz - 1, ny - 1)
    zid = receiver_row(mesh, rxLoc)
    F01 = F2d[zid:zid + 2, :].T.copy()
    sigma1 = np.asarray(sigma)[zid * ny:(zid + 1) * ny]
    zLen1 = zLen[zid]
    field = F2d.reshape(-1).copy()
    if isTE:
        Exr, Hyr = compFieldsAtRxTE(omega, rxLoc, yNode, zLen1, sigma1, F01)
        resp = _resp(omega, Exr, Hyr, dataType)
    else:
        Eyr, Hxr = compFieldsAtRxTM(omega, rxLoc, yNode, zLen1, sigma1, F01)
        resp = _resp(omega, Eyr, Hxr, dataType)
    return resp, field, Ainv


def default_factor(A):
    """Stand-in for `lu(Aii)` (UMFPACK) / `factorMUMPS(Aii,1)`: SciPy SuperLU."""
    return spla.splu(A.tocsc())


def assemble_mode(mesh: TensorMesh2D, isTE: bool, ii, io) -> CoeffMat:
    """Operator assembly, MT2DFwdSolver.jl:123-135 (TE) / :149-161 (TM)."""
    if not mesh.setup:
        setupTensorMesh2D(mesh)
    ncell = len(mesh.sigma)
    mu = MU0 * np.ones(ncell)
    F, Grad, AveCN, AveCF = mesh.Face, mesh.Grad, mesh.AveCN, mesh.AveCF
    sigma = np.asarray(mesh.sigma, dtype=np.float64)
    if isTE:
        Mnode = ops.sdiag(AveCN @ (F @ sigma))
        Medge = ops.sdiag(AveCF @ (F @ (1.0 / mu)))
    else:
        Mnode = ops.sdiag(AveCN @ (F @ mu))
        Medge = ops.sdiag(AveCF @ (F @ (1.0 / sigma)))
    dGrad = (Grad.T @ Medge @ Grad).tocsr()
    Mnode = Mnode.tocsr()
    return CoeffMat(dGrad[ii][:, ii].tocsc(), Mnode[ii][:, ii].tocsc(),
                    dGrad[ii][:, io].tocsc(), Mnode[ii][:, io].tocsc())


def MT2DFwdSolver(mesh: TensorMesh2D, data: MTData, factor_fn=default_factor):
    """`MT2DFwdSolver` MT2DFwdSolver.jl:74-216 -> (predData, MT2DFwdData)."""
    if not mesh.setup:
        setupTensorMesh2D(mesh)
    ny, nz = len(mesh.yLen), len(mesh.zLen)
    freqs, rxLoc = data.freqs, data.rxLoc
    nFreq, nRx = len(freqs), rxLoc.shape[0]
    nNode = (ny + 1) * (nz + 1)
    ii, io = ops.getBoundaryIndex(ny, nz)
    exte = np.zeros((nNode, nFreq), dtype=np.complex128)
    hxtm = np.zeros((nNode, nFreq), dtype=np.complex128)
    AinvTE, AinvTM = [None] * nFreq, [None] * nFreq
    respTE = respTM = None
    if data.compTE:
        coe = assemble_mode(mesh, True, ii, io)
        respTE = np.zeros((nFreq * nRx, 2))
        for j in range(nFreq):
            respTE[j * nRx:(j + 1) * nRx, :], exte[:, j], AinvTE[j] = _solve_mode(
                freqs[j], mesh, coe, rxLoc, data.dataType, True, factor_fn)
    if data.compTM:
        coe = assemble_mode(mesh, False, ii, io)
        respTM = np.zeros((nFreq * nRx, 2))
        for j in range(nFreq):
            respTM[j * nRx:(j + 1) * nRx, :], hxtm[:, j], AinvTM[j] = _solve_mode(
                freqs[j], mesh, coe, rxLoc, data.dataType, False, factor_fn)
    if "Impedance" in data.dataType:
        if data.compTE and not data.compTM:
            pred = respTE[:, 0] + 1j * respTE[:, 1]
        elif data.compTM and not data.compTE:
            pred = respTM[:, 0] + 1j * respTM[:, 1]
        else:
            pte = respTE[:, 0] + 1j * respTE[:, 1]
            ptm = respTM[:, 0] + 1j * respTM[:, 1]
            pred = np.stack([pte, ptm], axis=1).reshape(-1)     # vec(transpose(hcat)) :183-187
    elif "Rho_Pha" in data.dataType:
        if data.compTE and not data.compTM:
            pred = respTE.reshape(-1)
        elif data.compTM and not data.compTE:
            pred = respTM.reshape(-1)
        else:
            pred = np.concatenate([respTE, respTM], axis=1).reshape(-1)
    else:
        raise ValueError(data.dataType)
    pred = pred[data.dataID]
    return pred, MT2DFwdData(exte, hxtm, AinvTE, AinvTM, "")
