"""Explicit Jacobian — restates the Impedance branch of `HMCMT/src/MTSensitivity/compJacMat.jl`
(:150-330; P, B, L, Q definitions :197-206, :271-281).  Test infrastructure: it is a *second* reference
source for the same derivative, used to pin `compJacTMatVec` (J^T v must equal (explicit J)^T v)."""
from __future__ import annotations

import numpy as np

from . import operators as ops
from .fileio import setupTensorMesh2D
from .forward import MU0
from .sensitivity import getBCDerivMatrix, getDataFuncSensTE, getDataFuncSensTM, preSetRxFieldSens


def compJacMat(exTE, hxTM, mesh, data, activeCell, AinvTE, AinvTM):
    """-> complex J (nData x nAC), rows in (freq, rx, comp) order for fully populated Impedance data."""
    yLen, zLen, origin = mesh.yLen, mesh.zLen, mesh.origin
    sigma = np.asarray(mesh.sigma, dtype=np.float64)
    ny, nz = len(yLen), len(zLen)
    if not mesh.setup:
        setupTensorMesh2D(mesh)
    F, Grad, AveCN, AveCF = mesh.Face, mesh.Grad, mesh.AveCN, mesh.AveCF
    nCell, nNode = ny * nz, (ny + 1) * (nz + 1)
    nAC = activeCell.shape[1]
    mu = MU0 * np.ones(nCell)
    ii, io = ops.getBoundaryIndex(ny, nz)
    yNode = np.concatenate([[0.0], np.cumsum(yLen)]) - origin[0]
    zNode = np.concatenate([[0.0], np.cumsum(zLen)]) - origin[1]
    rxs = preSetRxFieldSens(data.rxLoc, yNode, zNode, sigma)
    zid = rxs.zid
    id0 = slice(zid * (ny + 1), (zid + 1) * (ny + 1))
    id1 = slice((zid + 1) * (ny + 1), (zid + 2) * (ny + 1))
    dGradTE = (Grad.T @ ops.sdiag(AveCF @ (F @ (1.0 / mu))) @ Grad).tocsr()
    MsigCN = ops.sdiag(AveCN @ (F @ sigma)).tocsr()
    dMsigCN = (AveCN[ii, :] @ F @ activeCell).tocsr()
    dGradTM = (Grad.T @ ops.sdiag(AveCF @ (F @ (1.0 / sigma))) @ Grad).tocsr()
    MmuCN = ops.sdiag(AveCN @ (F @ mu)).tocsr()
    Gradii, Gradio = Grad[:, ii], Grad[:, io]
    dMsigF = (AveCF @ F @ ops.sdiag(-1.0 / sigma ** 2) @ activeCell).tocsr()
    nRx, nFreq = data.rxLoc.shape[0], len(data.freqs)
    Acell = activeCell.toarray()
    rows = []
    for iF in range(nFreq):
        freq = data.freqs[iF]
        omega = 2 * np.pi * freq
        blocks = {}
        if data.compTE:
            AioTE = (dGradTE[ii][:, io] + 1j * omega * MsigCN[ii][:, io]).tocsr()
            dBC, _ = getBCDerivMatrix(freq, yLen, zLen, sigma, "E")
            dBC = dBC @ Acell
            PplusB = -1j * omega * (ops.sdiag(exTE[ii, iF]) @ dMsigCN).toarray() - AioTE @ dBC
            dE = np.zeros((nNode, nAC), dtype=np.complex128)
            dE[ii, :] = AinvTE[iF].solve(np.asarray(PplusB))
            dE[io, :] = dBC
            Ex01 = np.stack([exTE[id0, iF], exTE[id1, iF]], axis=1)
            L, Q = getDataFuncSensTE(omega, rxs, Ex01)
            blocks["TE"] = L @ dE + (Q @ activeCell).toarray()
        if data.compTM:
            AioTM = (dGradTM[ii][:, io] + 1j * omega * MmuCN[ii][:, io]).tocsr()
            dBC, bc = getBCDerivMatrix(freq, yLen, zLen, sigma, "H")
            dBC = dBC @ Acell
            PplusB = (-(Gradii.T @ ops.sdiag(Gradii @ hxTM[ii, iF]) @ dMsigF).toarray() - AioTM @ dBC
                      - (Gradii.T @ ops.sdiag(Gradio @ bc) @ dMsigF).toarray())
            dH = np.zeros((nNode, nAC), dtype=np.complex128)
            dH[ii, :] = AinvTM[iF].solve(np.asarray(PplusB))
            dH[io, :] = dBC
            Hx01 = np.stack([hxTM[id0, iF], hxTM[id1, iF]], axis=1)
            L, Q = getDataFuncSensTM(omega, rxs, Hx01)
            blocks["TM"] = L @ dH + (Q @ activeCell).toarray()
        for r in range(nRx):                       # (freq, rx, comp) interleave: MT2DFwdSolver.jl:183-187
            if "TE" in blocks:
                rows.append(blocks["TE"][r])
            if "TM" in blocks:
                rows.append(blocks["TM"][r])
    J = np.asarray(rows)
    return J[data.dataID[: J.shape[0]]] if J.shape[0] == len(data.dataID) else J
