"""Mesh operators — restates `HMCMT/src/MTFwdSolver/MT2DOperators.jl` (test infrastructure).

All matrices are scipy.sparse CSR/CSC float64.  Cell index c = k*ny + j (y fastest),
node index n = k*(ny+1) + j, 0-based (reference: 1-based).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def spunit(n: int) -> sp.csr_matrix:
    """`spunit` MT2DOperators.jl:140-142."""
    return sp.identity(n, dtype=np.float64, format="csr")


def sdiag(v) -> sp.csr_matrix:
    """`sdiag` MT2DOperators.jl:198-200."""
    return sp.diags(np.asarray(v), 0, format="csr")


def ddx(n: int) -> sp.csr_matrix:
    """1-D difference, node -> centre, n x (n+1).  `ddx` MT2DOperators.jl:161-163."""
    return sp.diags([-np.ones(n), np.ones(n)], [0, 1], shape=(n, n + 1), format="csr")


def av(n: int) -> sp.csr_matrix:
    """1-D average node -> centre, n x (n+1).  `av` MT2DOperators.jl:172-174;
    identical to `avnc` HMCUtility.jl:87-91."""
    return sp.diags([0.5 * np.ones(n), 0.5 * np.ones(n)], [0, 1], shape=(n, n + 1), format="csr")


avnc = av


def avcn(n: int) -> sp.csr_matrix:
    """1-D average centre -> node, (n+1) x n, boundary weights 1.0.
    `avcn` MT2DOperators.jl:183-190."""
    a = sp.diags([0.5 * np.ones(n), 0.5 * np.ones(n)], [-1, 0], shape=(n + 1, n), format="lil")
    a[0, 0] = 1.0
    a[n, n - 1] = 1.0
    return a.tocsr()


def meshGeoFace2D(d1, d2) -> sp.csr_matrix:
    """Cell areas on the diagonal.  MT2DOperators.jl:84-88."""
    return sp.kron(sdiag(d2), sdiag(d1), format="csr")


def meshGeoEdgeInv2D(d1, d2) -> sp.csr_matrix:
    """Inverse edge lengths: y-edges first, then z-edges.  MT2DOperators.jl:105-115."""
    n1, n2 = len(d1), len(d2)
    L1 = sp.kron(spunit(n2 + 1), sdiag(1.0 / np.asarray(d1)))
    L2 = sp.kron(sdiag(1.0 / np.asarray(d2)), spunit(n1 + 1))
    return sp.block_diag([L1, L2], format="csr")


def getNodalGradient2D(d1, d2) -> sp.csr_matrix:
    """Nodal gradient (edges x nodes).  MT2DOperators.jl:35-48."""
    n1, n2 = len(d1), len(d2)
    G1 = sp.kron(spunit(n2 + 1), ddx(n1))
    G2 = sp.kron(ddx(n2), spunit(n1 + 1))
    Grad = sp.vstack([G1, G2], format="csr")
    return (meshGeoEdgeInv2D(d1, d2) @ Grad).tocsr()


def getCellGradient2D(d1, d2) -> sp.csr_matrix:
    """Unscaled +-1 cell differences.  MT2DOperators.jl:52-63."""
    n1, n2 = len(d1), len(d2)
    G1 = sp.kron(spunit(n2), ddx(n1 - 1))
    G2 = sp.kron(ddx(n2 - 1), spunit(n1))
    return sp.vstack([G1, G2], format="csr")


def aveCell2Node2D(n) -> sp.csr_matrix:
    """MT2DOperators.jl:118-122."""
    return sp.kron(avcn(n[1]), avcn(n[0]), format="csr")


def aveCell2Face2D(n) -> sp.csr_matrix:
    """[A2; A1] — matches the edge ordering of Grad.  MT2DOperators.jl:126-130."""
    A1 = sp.kron(spunit(n[1]), avcn(n[0]))
    A2 = sp.kron(avcn(n[1]), spunit(n[0]))
    return sp.vstack([A2, A1], format="csr")


def getBoundaryIndex(ny: int, nz: int):
    """Interior / boundary node index lists (0-based).  MT2DFwdSolver.jl:227-248.

    ii: interior nodes, y fastest; io = [top | left | right | bottom].
    """
    nNode = (ny + 1) * (nz + 1)
    idx2D = np.arange(nNode).reshape(nz + 1, ny + 1)       # idx2D[k, j]
    ii = idx2D[1:-1, 1:-1].reshape(-1)
    it = idx2D[0, :]
    il = idx2D[1:, 0]
    ir = idx2D[1:, -1]
    ib = idx2D[-1, 1:-1]
    io = np.concatenate([it, il, ir, ib])
    return ii, io
