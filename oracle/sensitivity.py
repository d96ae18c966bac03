"""Adjoint gradient J^T v — restates `HMCMT/src/MTSensitivity/{compJacTMatVec,dataFuncSens,
sensUtils,MT1DSensitivity}.jl` (test infrastructure).

Only `DataType: Impedance` is restated: the reference's gradient code tests the string
"Rho_Phs" while its reader/forward use "Rho_Pha", so apparent-resistivity data cannot reach
the gradient there.
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np
import scipy.sparse as sp

from . import operators as ops
from .fileio import TensorMesh2D, MTData, setupTensorMesh2D
from .forward import MU0

# ----------------------------------------------------------------------------------------
# sensUtils.jl


def linearInterp(point, x):
    """`linearInterp` sensUtils.jl:133-161 (returns 0-based indices)."""
    ind = int(np.argmin(np.abs(point - x)))
    if point - x[ind] > 0:
        indL, indR = ind, ind + 1
    else:
        indL, indR = ind - 1, ind
    n = len(x)
    indL = max(min(indL, n - 1), 0)
    indR = max(min(indR, n - 1), 0)
    if indL == indR:
        return indL, indR, 0.5, 0.5
    xLen = x[indR] - x[indL]
    wL = 1 - (point - x[indL]) / xLen
    wR = 1 - (x[indR] - point) / xLen
    return indL, indR, wL, wR


def linearInterpMat(points, x):
    """`linearInterpMat` sensUtils.jl:63-83 -> (len(x) x npts) sparse.

    A column assigned through `sparsevec([indL;indR],[wL;wR])` sums duplicate indices, so the
    degenerate indL==indR case yields a single weight 1.0.
    """
    rows, cols, vals = [], [], []
    for i, p in enumerate(points):
        iL, iR, wL, wR = linearInterp(p, x)
        rows += [iL, iR]
        cols += [i, i]
        vals += [wL, wR]
    return sp.csr_matrix((vals, (rows, cols)), shape=(len(x), len(points)))


@dataclass
class PreRxSens:
    """`PreRxSens` MTSensitivity.jl:20-47 (zid 0-based)."""
    zid: int
    dFn0: sp.csr_matrix
    dFn1: sp.csr_matrix
    sigma1: np.ndarray
    dsigma1: sp.csr_matrix
    yLen: np.ndarray
    zLen1: float
    linRxMap: sp.csr_matrix
    linRxMap2: sp.csr_matrix


def preSetRxFieldSens(rxLoc, yNode, zNode, sigma) -> PreRxSens:
    """`preSetRxFieldSens` sensUtils.jl:17-52."""
    ny, nz = len(yNode) - 1, len(zNode) - 1
    nNode, nCell = (ny + 1) * (nz + 1), ny * nz
    zLen = np.diff(zNode)
    zid = int(np.nonzero(np.abs(zNode - rxLoc[0, 1]) < 0.1)[0][0])
    Inode = ops.spunit(nNode)
    dFn0 = Inode[zid * (ny + 1):(zid + 1) * (ny + 1), :]
    dFn1 = Inode[(zid + 1) * (ny + 1):(zid + 2) * (ny + 1), :]
    Icell = ops.spunit(nCell)
    sigma1 = np.asarray(sigma)[zid * ny:(zid + 1) * ny]
    dsigma1 = Icell[zid * ny:(zid + 1) * ny, :]
    yLen = np.diff(yNode)
    linRxMap = linearInterpMat(rxLoc[:, 0], yNode)
    yCen = (yNode[:-1] + yNode[1:]) / 2.0
    linRxMap2 = linearInterpMat(rxLoc[:, 0], yCen)
    return PreRxSens(zid, dFn0, dFn1, sigma1, dsigma1, yLen, zLen[zid], linRxMap, linRxMap2)


def _copy_ends(M):
    """dX[1,:] = dX[2,:]; dX[end,:] = dX[end-1,:] on a (ny+1) x n sparse matrix whose first and
    last rows are empty (dataFuncSens.jl:84-86)."""
    M = M.tolil()
    M[0, :] = M[1, :]
    M[-1, :] = M[-2, :]
    return M.tocsr()


def _pad_rows(inner, ny):
    """Place an (ny-1) x n matrix into rows 1..ny-1 of an (ny+1) x n zero matrix."""
    n = inner.shape[1]
    z = sp.csr_matrix((1, n), dtype=np.complex128)
    return sp.vstack([z, inner.astype(np.complex128), z], format="csr")


# ----------------------------------------------------------------------------------------
# dataFuncSens.jl


def getDataFuncSensTE(omega, rx: PreRxSens, Ex01):
    """`getDataFuncSensTE` dataFuncSens.jl:21-176 (Impedance branch) -> L (nRx x nNode), Q (nRx x nCell)."""
    dEx0, dEx1 = rx.dFn0, rx.dFn1
    sigma1, dsigma1, yLen, zLen1 = rx.sigma1, rx.dsigma1, rx.yLen, rx.zLen1
    ny = len(yLen)
    mu = MU0 * np.ones(ny)
    Bz0 = (ops.ddx(ny) @ Ex01[:, 0]) / yLen / (1j * omega)
    dtmp = ops.sdiag(1.0 / yLen / (1j * omega)) @ ops.ddx(ny)
    dBz0 = dtmp @ dEx0
    dBz1 = dtmp @ dEx1
    dHzQ = ops.sdiag(1.0 / mu) @ (0.75 * dBz0 + 0.25 * dBz1)
    dHyH = -(dEx1[1:-1, :] - dEx0[1:-1, :]) / zLen1 / (1j * omega * MU0)
    ExQ = 0.75 * Ex01[1:-1, 0] + 0.25 * Ex01[1:-1, 1]
    dExQ = 0.75 * dEx0[1:-1, :] + 0.25 * dEx1[1:-1, :]
    avm = ops.avnc(ny - 1)
    ybar = avm @ yLen
    sigma1v = (avm @ (sigma1 * yLen)) / ybar
    dsigma1v = ops.sdiag(1.0 / ybar) @ avm @ ops.sdiag(yLen) @ dsigma1
    ddHzQ = ops.sdiag(1.0 / ybar) @ ops.ddx(ny - 1) @ dHzQ
    # forward quantities (needed for Z)
    Bz1 = (ops.ddx(ny) @ Ex01[:, 1]) / yLen / (1j * omega)
    HzQ = (0.75 * Bz0 + 0.25 * Bz1) / mu
    HyH = -(Ex01[1:-1, 1] - Ex01[1:-1, 0]) / zLen1 / (1j * omega * MU0)
    dHzQ_dy = (ops.ddx(ny - 1) @ HzQ) / ybar
    Hy0 = np.zeros(ny + 1, dtype=np.complex128)
    Hy0[1:-1] = HyH - (dHzQ_dy - sigma1v * ExQ) * (0.5 * zLen1)
    Hy0[0], Hy0[-1] = Hy0[1], Hy0[-2]
    dHy0 = _copy_ends(_pad_rows(dHyH - (ddHzQ - ops.sdiag(sigma1v) @ dExQ) * (0.5 * zLen1), ny))
    dHy0_dsig = _copy_ends(_pad_rows(0.5 * zLen1 * (ops.sdiag(ExQ) @ dsigma1v), ny))
    Exr = rx.linRxMap.T @ Ex01[:, 0]
    Hyr = rx.linRxMap.T @ Hy0
    dExr = rx.linRxMap.T @ dEx0
    dHyr = rx.linRxMap.T @ dHy0
    dHyr_dsig = rx.linRxMap.T @ dHy0_dsig
    dZ = ops.sdiag(1.0 / Hyr) @ dExr - ops.sdiag(Exr / Hyr ** 2) @ dHyr
    dZ_dsig = -ops.sdiag(Exr / Hyr ** 2) @ dHyr_dsig
    return dZ.tocsr(), dZ_dsig.tocsr()


def getDataFuncSensTM(omega, rx: PreRxSens, Hx01):
    """`getDataFuncSensTM` dataFuncSens.jl:197-344 (Impedance branch)."""
    dHx0, dHx1 = rx.dFn0, rx.dFn1
    sigma1, dsigma1, yLen, zLen1 = rx.sigma1, rx.dsigma1, rx.yLen, rx.zLen1
    ny = len(yLen)
    Jz0 = -(ops.ddx(ny) @ Hx01[:, 0]) / yLen
    Jz1 = -(ops.ddx(ny) @ Hx01[:, 1]) / yLen
    dtmp = -ops.sdiag(1.0 / yLen) @ ops.ddx(ny)
    dJz0 = dtmp @ dHx0
    dJz1 = dtmp @ dHx1
    EzQ = (0.75 * Jz0 + 0.25 * Jz1) / sigma1
    dEzQ = ops.sdiag(1.0 / sigma1) @ (0.75 * dJz0 + 0.25 * dJz1)
    dEzQ_dsig = ops.sdiag(0.75 * Jz0 + 0.25 * Jz1) @ ops.sdiag(-1.0 / sigma1 ** 2) @ dsigma1
    JyH = (Hx01[1:-1, 1] - Hx01[1:-1, 0]) / zLen1
    avm = ops.avnc(ny - 1)
    ybar = avm @ yLen
    rho1v = (avm @ ((1.0 / sigma1) * yLen)) / ybar
    dJyH = (dHx1[1:-1, :] - dHx0[1:-1, :]) / zLen1
    EyH = JyH * rho1v
    dEyH = ops.sdiag(rho1v) @ dJyH
    drho1v = ops.sdiag(1.0 / ybar) @ avm @ ops.sdiag(yLen) @ ops.sdiag(-1.0 / sigma1 ** 2) @ dsigma1
    dEyH_dsig = ops.sdiag(JyH) @ drho1v
    HxQ = 0.75 * Hx01[1:-1, 0] + 0.25 * Hx01[1:-1, 1]
    dHxQ = 0.75 * dHx0[1:-1, :] + 0.25 * dHx1[1:-1, :]
    dEzQ_dy = (ops.ddx(ny - 1) @ EzQ) / ybar
    dtmp = ops.sdiag(1.0 / ybar) @ ops.ddx(ny - 1)
    ddEzQ = dtmp @ dEzQ
    ddEzQ_dsig = dtmp @ dEzQ_dsig
    Ey0 = np.zeros(ny + 1, dtype=np.complex128)
    Ey0[1:-1] = EyH - (dEzQ_dy + 1j * omega * MU0 * HxQ) * (0.5 * zLen1)
    Ey0[0], Ey0[-1] = Ey0[1], Ey0[-2]
    dEy0 = _copy_ends(_pad_rows(dEyH - (ddEzQ + 1j * omega * MU0 * dHxQ) * (0.5 * zLen1), ny))
    dEy0_dsig = _copy_ends(_pad_rows(dEyH_dsig - ddEzQ_dsig * (0.5 * zLen1), ny))
    Hxr = rx.linRxMap.T @ Hx01[:, 0]
    Eyr = rx.linRxMap.T @ Ey0
    dHxr = rx.linRxMap.T @ dHx0
    dEyr = rx.linRxMap.T @ dEy0
    dEyr_dsig = rx.linRxMap.T @ dEy0_dsig
    dZ = ops.sdiag(1.0 / Hxr) @ dEyr - ops.sdiag(Eyr / Hxr ** 2) @ dHxr
    dZ_dsig = ops.sdiag(1.0 / Hxr) @ dEyr_dsig
    return dZ.tocsr(), dZ_dsig.tocsr()


# ----------------------------------------------------------------------------------------
# MT1DSensitivity.jl


def compImpJacMatrix(freq, sig1d, thick1d):
    """`compImpJacMatrix` MT1DSensitivity.jl:188-243.  sig1d includes the half-space."""
    nLayer = len(sig1d)
    omega = 2 * np.pi * freq
    iom = 1j * omega * MU0
    Z = 0j
    dZ_ZP1 = np.zeros(nLayer, dtype=np.complex128)
    dZ_sigma = np.zeros(nLayer, dtype=np.complex128)
    zimpDeri = np.zeros(nLayer, dtype=np.complex128)
    for j in range(nLayer - 1, -1, -1):
        k = np.sqrt(-iom * sig1d[j])
        Zt = omega * MU0 / k
        dZt = 1j * (omega * MU0) ** 2 / (2 * k ** 3)
        if j == nLayer - 1:
            Z = Zt
            dZ_sigma[j] = dZt
            continue
        RI = (Zt - Z) / (Zt + Z)
        theEXP = np.exp(-2j * k * thick1d[j])
        L = RI * theEXP
        Ztmp = Zt * (1 - L) / (1 + L)
        dL = 2 * Z / (Zt + Z) ** 2 * theEXP * dZt + (-2j * thick1d[j] * L) * (-iom / 2 / k)
        dZ_ZP1[j] = 4 * Zt * Zt * theEXP / ((1 + L) * (Zt + Z)) ** 2
        dZ_sigma[j] = dZt * (1 - L) / (1 + L) + Zt * (-2) / (1 + L) ** 2 * dL
        Z = Ztmp
    for iLayer in range(nLayer - 1, 0, -1):
        dZ_ZPN = 1.0 + 0j
        for j in range(iLayer):
            dZ_ZPN = dZ_ZPN * dZ_ZP1[j]
        zimpDeri[iLayer] = dZ_ZPN * dZ_sigma[iLayer]
    zimpDeri[0] = dZ_sigma[0]
    return Z, zimpDeri


def mt1DFieldSensMatrix(freq, sig1d, zNode, source="E", fTop=1.0):
    """`mt1DFieldSensMatrix` MT1DSensitivity.jl:25-176 -> (field[nz+1], dF[(nz+1) x nz]).

    Quirks kept: no eps0 term in the wavenumber (:59), the half-space column is dropped
    (:162-164), and the overflow guard zeroes only the lower-right block (:144-151).
    """
    sig1d = np.asarray(sig1d, dtype=np.float64)
    omega = 2 * np.pi * freq
    omu = omega * MU0
    sigma = np.concatenate([sig1d, sig1d[-1:]])
    nLayer = len(sigma)
    zLen = np.diff(np.asarray(zNode, dtype=np.float64))
    z1, dz1 = compImpJacMatrix(freq, sigma, zLen)
    eLayer = np.zeros((2, nLayer), dtype=np.complex128)
    dEu = np.zeros((nLayer, nLayer), dtype=np.complex128)
    dEd = np.zeros((nLayer, nLayer), dtype=np.complex128)
    dHu = np.zeros((nLayer, nLayer), dtype=np.complex128)
    dHd = np.zeros((nLayer, nLayer), dtype=np.complex128)
    ka = np.sqrt(-1j * omu * sigma)
    dkaVec = (-1j * omu / 2) / ka
    dka = np.diag(dkaVec)
    k1 = ka[0]
    if source == "E":
        eLayer[0, 0] = 0.5 * fTop * (1 - omu / (z1 * k1))
        eLayer[1, 0] = 0.5 * fTop * (1 + omu / (z1 * k1))
        dEu[0, :] = 0.5 * fTop * omu / (z1 * k1) * (1 / z1 * dz1 + 1 / k1 * dka[0, :])
        dEd[0, :] = -dEu[0, :]
        dHu[0, :] = -eLayer[0, 0] / omu * dka[0, :] - ka[0] / omu * dEu[0, :]
        dHd[0, :] = eLayer[1, 0] / omu * dka[0, :] + ka[0] / omu * dEd[0, :]
    elif source == "H":
        hu = 0.5 * fTop * (1 - z1 * k1 / omu)
        hd = 0.5 * fTop * (1 + z1 * k1 / omu)
        eLayer[0, 0] = -omu / k1 * hu
        eLayer[1, 0] = omu / k1 * hd
        dHu[0, :] = -0.5 * fTop / omu * (z1 * dka[0, :] + k1 * dz1)
        dHd[0, :] = -dHu[0, :]
        dEu[0, :] = 0.5 * fTop * (dz1 + (omu / k1 ** 2) * dka[0, :])
        dEd[0, :] = 0.5 * fTop * (dz1 - (omu / k1 ** 2) * dka[0, :])
    else:
        raise ValueError(source)
    with np.errstate(over="ignore", invalid="ignore"):
        expt = np.exp(1j * ka[:-1] * zLen)
        expr = 1.0 / expt
        dexptv = 1j * zLen * expt * dkaVec[:-1]
        dexprv = -1j * zLen * expr * dkaVec[:-1]
        dexpt = np.zeros((nLayer - 1, nLayer), dtype=np.complex128)
        dexpr = np.zeros((nLayer - 1, nLayer), dtype=np.complex128)
        dexpt[np.arange(nLayer - 1), np.arange(nLayer - 1)] = dexptv
        dexpr[np.arange(nLayer - 1), np.arange(nLayer - 1)] = dexprv
        kr = ka[:-1] / ka[1:]
        dkr = np.zeros((nLayer - 1, nLayer), dtype=np.complex128)
        for j in range(nLayer - 1):
            dkr[j, :] = dka[j, :] / ka[j + 1] - ka[j] / (ka[j + 1] ** 2) * dka[j + 1, :]
        mix11 = (1 + kr) * expt
        mix12 = (1 - kr) * expr
        mix21 = (1 - kr) * expt
        mix22 = (1 + kr) * expr
        dmix11 = (1 + kr)[:, None] * dexpt + expt[:, None] * dkr
        dmix12 = (1 - kr)[:, None] * dexpr - expr[:, None] * dkr
        dmix21 = (1 - kr)[:, None] * dexpt - expt[:, None] * dkr
        dmix22 = (1 + kr)[:, None] * dexpr + expr[:, None] * dkr
        for j in range(nLayer - 1):
            eu, ed = eLayer[0, j], eLayer[1, j]
            eLayer[0, j + 1] = 0.5 * ((1 + kr[j]) * expt[j] * eu + (1 - kr[j]) * expr[j] * ed)
            eLayer[1, j + 1] = 0.5 * ((1 - kr[j]) * expt[j] * eu + (1 + kr[j]) * expr[j] * ed)
            dEu[j + 1, :] = 0.5 * (dmix11[j, :] * eu + mix11[j] * dEu[j, :] + dmix12[j, :] * ed + mix12[j] * dEd[j, :])
            dEd[j + 1, :] = 0.5 * (dmix21[j, :] * eu + mix21[j] * dEu[j, :] + dmix22[j, :] * ed + mix22[j] * dEd[j, :])
            epu, epd = eLayer[0, j + 1], eLayer[1, j + 1]
            dHu[j + 1, :] = -epu / omu * dka[j + 1, :] - ka[j + 1] / omu * dEu[j + 1, :]
            dHd[j + 1, :] = epd / omu * dka[j + 1, :] + ka[j + 1] / omu * dEd[j + 1, :]
            e2 = abs(eLayer[0, j + 1] + eLayer[1, j + 1])
            e1 = abs(eLayer[0, j] + eLayer[1, j])
            if e2 - e1 > 0.0 or np.isnan(e2):
                eLayer[:, j + 1:] = 0.0
                dEu[j + 1:, j + 1:] = 0.0
                dEd[j + 1:, j + 1:] = 0.0
                dHu[j + 1:, j + 1:] = 0.0
                dHd[j + 1:, j + 1:] = 0.0
                break
    dE = (dEu + dEd)[:, :-1]
    dH = (dHu + dHd)[:, :-1]
    if source == "E":
        return eLayer.sum(axis=0), dE
    hField = (-ka * eLayer[0] + ka * eLayer[1]) / omu
    return hField, dH


def bc_profiles(freq, yLen, zLen, sigma, source):
    """The three 1-D sensitivity solves `getBCDerivMatrix` is built from
    (MT1DSensitivity.jl:274-286, 313-314): left column, right column, row-mean profile."""
    ny, nz = len(yLen), len(zLen)
    zNode = np.concatenate([[0.0], np.cumsum(zLen)])
    sig2D = np.asarray(sigma).reshape(nz, ny)
    left = mt1DFieldSensMatrix(freq, sig2D[:, 0], zNode, source, 1.0)
    right = mt1DFieldSensMatrix(freq, sig2D[:, -1], zNode, source, 1.0)
    mean = mt1DFieldSensMatrix(freq, sig2D.mean(axis=1), zNode, source, 1.0)
    return left, right, mean


def getBCDerivMatrix(freq, yLen, zLen, sigma, source):
    """`getBCDerivMatrix` MT1DSensitivity.jl:253-333: dense (nb x nCell) dBC and bc.  Literal
    (small meshes only: the reference materialises 2(ny+nz) x nCell complex)."""
    ny, nz = len(yLen), len(zLen)
    nb, ncell = 2 * (ny + nz), ny * nz
    bc = np.zeros(nb, dtype=np.complex128)
    dBC = np.zeros((nb, ncell), dtype=np.complex128)
    bc[0:ny + 1] = 1.0
    (ebL, dEL), (ebR, dER), (ebM, dEM) = bc_profiles(freq, yLen, zLen, sigma, source)
    idx = np.arange(ny + 1, ny + nz + 1)
    bc[idx] = ebL[1:]
    dBC[np.ix_(idx, np.arange(0, ncell, ny))] = dEL[1:, :]
    idx = np.arange(ny + nz + 1, ny + 2 * nz + 1)
    bc[idx] = ebR[1:]
    dBC[np.ix_(idx, np.arange(ny - 1, ncell, ny))] = dER[1:, :]
    for j in range(1, ny):          # reference j = 2..ny
        y1, y2 = yLen[j - 1], yLen[j]
        row = ny + 2 * nz + j
        bc[row] = ebM[-1]
        dBC[row, np.arange(j - 1, ncell, ny)] += dEM[-1, :] * y1 / (y1 + y2)
        dBC[row, np.arange(j, ncell, ny)] += dEM[-1, :] * y2 / (y1 + y2)
    return dBC, bc


def bc_deriv_apply_T(freq, yLen, zLen, sigma, source, t):
    """Matrix-free transpose(dBC) @ t and the derivative routine's bc — numerically the same
    contraction as `transpose(dBC) * t` (compJacTMatVec.jl:240-243) without the dense matrix."""
    ny, nz = len(yLen), len(zLen)
    nb, ncell = 2 * (ny + nz), ny * nz
    (ebL, dEL), (ebR, dER), (ebM, dEM) = bc_profiles(freq, yLen, zLen, sigma, source)
    bc = np.zeros(nb, dtype=np.complex128)
    bc[0:ny + 1] = 1.0
    bc[ny + 1:ny + nz + 1] = ebL[1:]
    bc[ny + nz + 1:ny + 2 * nz + 1] = ebR[1:]
    bc[ny + 2 * nz + 1:] = ebM[-1]
    out = np.zeros((nz, ny), dtype=np.complex128)
    out[:, 0] += dEL[1:, :].T @ t[ny + 1:ny + nz + 1]
    out[:, -1] += dER[1:, :].T @ t[ny + nz + 1:ny + 2 * nz + 1]
    tb = t[ny + 2 * nz + 1:]                       # nodes j = 1..ny-1 (0-based)
    yl = np.asarray(yLen)
    w1 = yl[:-1] / (yl[:-1] + yl[1:])
    w2 = yl[1:] / (yl[:-1] + yl[1:])
    colw = np.zeros(ny, dtype=np.complex128)
    colw[:-1] += w1 * tb
    colw[1:] += w2 * tb
    out += dEM[-1, :][:, None] * colw[None, :]
    return out.reshape(-1), bc


# ----------------------------------------------------------------------------------------
# compJacTMatVec.jl


def compJacTMatVec(exTE, hxTM, datVec, mesh: TensorMesh2D, data: MTData, activeCell=None,
                   AinvTE=None, AinvTM=None, dense_bc=False, parts=None):
    """`compJacTMatVec` compJacTMatVec.jl:8-329 -> real(J^T v) over active cells.

    `dense_bc=True` follows the reference's dense `dBC` literally; False uses the matrix-free
    contraction (identical arithmetic per entry, different summation order inside the matvec).
    `parts`, if a dict, receives per-(mode,freq) intermediate vectors for kernel-level tests.
    """
    from .operators import getBoundaryIndex
    yLen, zLen, origin = mesh.yLen, mesh.zLen, mesh.origin
    sigma = np.asarray(mesh.sigma, dtype=np.float64)
    ny, nz = len(yLen), len(zLen)
    freqs, rxLoc = data.freqs, data.rxLoc
    yNode = np.concatenate([[0.0], np.cumsum(yLen)]) - origin[0]
    zNode = np.concatenate([[0.0], np.cumsum(zLen)]) - origin[1]
    nFreq, nRx = len(freqs), rxLoc.shape[0]
    nCell = ny * nz
    if activeCell is None:
        activeCell = ops.spunit(nCell)
    nAC = activeCell.shape[1]
    mu = MU0 * np.ones(nCell)
    if not mesh.setup:
        setupTensorMesh2D(mesh)
    F, Grad, AveCN, AveCF = mesh.Face, mesh.Grad, mesh.AveCN, mesh.AveCF
    ii, io = getBoundaryIndex(ny, nz)
    if data.compTE:
        MsigCN = ops.sdiag(AveCN @ (F @ sigma)).tocsr()
        dGradTE = (Grad.T @ ops.sdiag(AveCF @ (F @ (1.0 / mu))) @ Grad).tocsr()
        rAioTE = dGradTE[ii][:, io]
        iAioTE = MsigCN[ii][:, io]
        dMsigCN = (AveCN[ii, :] @ F @ activeCell).tocsr()
    if data.compTM:
        MmuCN = ops.sdiag(AveCN @ (F @ mu)).tocsr()
        dGradTM = (Grad.T @ ops.sdiag(AveCF @ (F @ (1.0 / sigma))) @ Grad).tocsr()
        rAioTM = dGradTM[ii][:, io]
        iAioTM = MmuCN[ii][:, io]
        Gradii = Grad[:, ii]
        Gradio = Grad[:, io]
        dMsigF = (AveCF @ F @ ops.sdiag(-1.0 / sigma ** 2) @ activeCell).tocsr()
    if "Impedance" not in data.dataType:
        raise NotImplementedError("only DataType Impedance reaches the gradient in the reference")
    iZXY = iZYX = 0
    for j, c in enumerate(data.dataComp):
        if c == "ZXY":
            iZXY = j + 1
        elif c == "ZYX":
            iZYX = j + 1
    rxs = preSetRxFieldSens(rxLoc, yNode, zNode, sigma)
    zid = rxs.zid
    id0 = slice(zid * (ny + 1), (zid + 1) * (ny + 1))
    id1 = slice((zid + 1) * (ny + 1), (zid + 2) * (ny + 1))
    JTv = np.zeros(nAC, dtype=np.complex128)
    QTv = np.zeros(nAC, dtype=np.complex128)
    datVec = np.asarray(datVec)
    for iFreq in range(nFreq):
        freq = freqs[iFreq]
        omega = 2 * np.pi * freq
        indF = np.nonzero(data.freqID == iFreq + 1)[0]
        if len(indF) == 0:
            continue
        subRxID = data.rxID[indF]
        subDcID = data.dtID[indF]
        datTmp = np.conj(datVec[indF])
        calTE = any("XY" in data.dataComp[d - 1] for d in subDcID)
        calTM = any("YX" in data.dataComp[d - 1] for d in subDcID)
        if calTE:
            Ex01 = np.stack([exTE[id0, iFreq], exTE[id1, iFreq]], axis=1)
            L, Q = getDataFuncSensTE(omega, rxs, Ex01)
            idd1 = np.nonzero(subDcID == iZXY)[0]
            idr1 = subRxID[idd1] - 1
            sVec = L[idr1, :].T @ datTmp[idd1]
            qv = activeCell.T @ (Q[idr1, :].T @ datTmp[idd1])
            QTv = QTv + qv
            AioTE = (rAioTE + 1j * omega * iAioTE).tocsr()
            eVal = AinvTE[iFreq].solve(sVec[ii])
            eVal_io = sVec[io]
            PTv = -1j * omega * (dMsigCN.T @ (exTE[ii, iFreq] * eVal))
            tvec = -(AioTE.T @ eVal) + eVal_io
            if dense_bc:
                dBC, _ = getBCDerivMatrix(freq, yLen, zLen, sigma, "E")
                dBCa = dBC @ activeCell
                BT = dBCa.T @ (-(AioTE.T @ eVal)) + dBCa.T @ eVal_io
            else:
                g, _ = bc_deriv_apply_T(freq, yLen, zLen, sigma, "E", tvec)
                BT = activeCell.T @ g
            JTv = JTv + PTv + BT
            if parts is not None:
                parts[("TE", iFreq)] = dict(s=sVec, lam=eVal, t=tvec, P=PTv, B=BT, q=qv)
        if calTM:
            Hx01 = np.stack([hxTM[id0, iFreq], hxTM[id1, iFreq]], axis=1)
            L, Q = getDataFuncSensTM(omega, rxs, Hx01)
            idd1 = np.nonzero(subDcID == iZYX)[0]
            idr1 = subRxID[idd1] - 1
            sVec = L[idr1, :].T @ datTmp[idd1]
            qv = activeCell.T @ (Q[idr1, :].T @ datTmp[idd1])
            QTv = QTv + qv
            AioTM = (rAioTM + 1j * omega * iAioTM).tocsr()
            eVal = AinvTM[iFreq].solve(sVec[ii])
            eVal_io = sVec[io]
            gl = -(Gradii @ eVal)
            PTv = dMsigF.T @ ((Gradii @ hxTM[ii, iFreq]) * gl)
            tvec = -(AioTM.T @ eVal) + eVal_io
            if dense_bc:
                dBC, bc = getBCDerivMatrix(freq, yLen, zLen, sigma, "H")
                dBCa = dBC @ activeCell
                BT = dBCa.T @ (-(AioTM.T @ eVal)) + dBCa.T @ eVal_io
            else:
                g, bc = bc_deriv_apply_T(freq, yLen, zLen, sigma, "H", tvec)
                BT = activeCell.T @ g
            BT2 = dMsigF.T @ ((Gradio @ bc) * gl)
            JTv = JTv + PTv + BT + BT2
            if parts is not None:
                parts[("TM", iFreq)] = dict(s=sVec, lam=eVal, t=tvec, P=PTv, B=BT, B2=BT2, q=qv, bc=bc)
    return np.real(JTv + QTv)
