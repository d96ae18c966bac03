"""Pins the CPU oracle (the reference ships no golden vectors for this path):
analytic half-space / layered-earth answers, and the two reference derivations of the Jacobian."""
import numpy as np
import pytest

from oracle import forward as ofwd
from oracle import jacobian as ojac
from oracle import operators as ops
from oracle import sensitivity as osens
from tests.helpers import load_example, tiny_problem

MU0 = 4e-7 * np.pi


def test_halfspace_impedance_dprism():
    """examples/dprism3d start model is a 0.01 S/m half-space: Z = sqrt(i w mu0 / sigma) up to mesh error
    (1.3e-3 at 100 Hz, 6.5e-3 at 1 Hz, 2e-2 at 0.01 Hz)."""
    mesh, data, inv, prior = load_example("dprism3d")
    pred, _ = ofwd.MT2DFwdSolver(mesh, data)
    p = pred.reshape(len(data.freqs), -1, 2)
    zan = np.sqrt(1j * 2 * np.pi * data.freqs * MU0 / 0.01)
    for i, tol in [(0, 2e-3), (5, 8e-3), (10, 2.5e-2)]:
        assert abs(p[i, 20, 0] - zan[i]) / abs(zan[i]) < tol        # ZXY
        assert abs(p[i, 20, 1] + zan[i]) / abs(zan[i]) < tol        # ZYX = -Z


def test_1d_field_matches_impedance_recursion():
    """E/H at the surface from mt1DAnalyticField equals mt1DImpedance (mt1DField.jl:23-98 vs :115-165)."""
    rng = np.random.default_rng(0)
    sig = np.exp(np.log(0.01) + rng.standard_normal(12))
    zn = np.concatenate([[0.0], np.cumsum(np.full(12, 150.0))])
    for f in [100.0, 1.0, 0.01]:
        e, h = ofwd.mt1DAnalyticField(f, sig, zn, True)
        z = ofwd.mt1DImpedance(np.array([f]), np.concatenate([sig, sig[-1:]]), zn)[0]
        assert abs(e[0] / h[0] - z) / abs(z) < 1e-12
        assert abs(e[0] - 1.0) < 1e-14


def test_layered_model_2d_matches_1d():
    """A laterally uniform layered model must give the 1-D impedance at every receiver."""
    mesh, data, inv, prior = load_example("dprism3d")
    ny, nz = mesh.gridSize
    lay = np.where(np.arange(nz) < 7, 1e-8, np.where(np.arange(nz) < 20, 0.02, 0.002))
    mesh.sigma = np.repeat(lay, ny)
    pred, _ = ofwd.MT2DFwdSolver(mesh, data)
    p = pred.reshape(len(data.freqs), -1, 2)
    zn = np.concatenate([[0.0], np.cumsum(mesh.zLen[7:])])
    z1 = ofwd.mt1DImpedance(data.freqs, np.concatenate([lay[7:], lay[-1:]]), zn)
    for i in [3, 6, 9]:
        assert abs(p[i, 10, 0] - z1[i]) / abs(z1[i]) < 3e-2
        assert np.abs(p[i, :, 0] - p[i, 10, 0]).max() / abs(z1[i]) < 1e-3      # (almost) no lateral variation


def test_dof_numbering_and_pattern():
    """nnz = 5N - 2(ny-1) - 2(nz-1), symmetric, imaginary part only on the diagonal."""
    mesh, data, inv, prior = tiny_problem()
    ny, nz = mesh.gridSize
    ii, io = ops.getBoundaryIndex(ny, nz)
    N = (ny - 1) * (nz - 1)
    assert len(ii) == N and len(io) == 2 * (ny + nz)
    assert ii[0] == (ny + 1) + 1 and ii[1] == ii[0] + 1              # y fastest
    coe = ofwd.assemble_mode(mesh, True, ii, io)
    A = (coe.rAii + 1j * 3.0 * coe.iAii).tocsc()
    assert A.nnz == 5 * N - 2 * (ny - 1) - 2 * (nz - 1)
    assert abs(A - A.T).max() == 0
    off = A - np.diag(A.diagonal())
    assert np.abs(np.imag(off)).max() == 0
    assert coe.rAio.nnz == 2 * (ny - 1) + 2 * (nz - 1)


def test_jtvec_equals_explicit_jacobian_transpose():
    """compJacTMatVec.jl (J^T v) against compJacMat.jl (explicit J): two reference derivations agree."""
    mesh, data, inv, prior = tiny_problem()
    pred, fwd = ofwd.MT2DFwdSolver(mesh, data)
    J = ojac.compJacMat(fwd.exTE, fwd.hxTM, mesh, data, inv.activeCell, fwd.AinvTE, fwd.AinvTM)
    rng = np.random.default_rng(5)
    v = rng.standard_normal(J.shape[0]) + 1j * rng.standard_normal(J.shape[0])
    g = osens.compJacTMatVec(fwd.exTE, fwd.hxTM, v, mesh, data, inv.activeCell, fwd.AinvTE, fwd.AinvTM, dense_bc=True)
    gJ = np.real(J.T @ np.conj(v))
    assert np.abs(g - gJ).max() / np.abs(gJ).max() < 1e-12
    g2 = osens.compJacTMatVec(fwd.exTE, fwd.hxTM, v, mesh, data, inv.activeCell, fwd.AinvTE, fwd.AinvTM, dense_bc=False)
    assert np.abs(g - g2).max() / np.abs(g).max() < 1e-13          # matrix-free dBC^T t == dense dBC


def test_gradient_vs_finite_differences_upper_cells():
    """Reference-style adjoint vs central FD on cells away from the bottom rows (the reference's own
    BC approximations spoil the deepest rows)."""
    from oracle import sampler as osamp
    mesh, data, inv, prior = tiny_problem()
    m0 = inv.strModel.copy()
    _, _, g = osamp.compDataGradient(mesh, data, inv, prior)
    ny = mesh.gridSize[0]
    h = 1e-5
    for cell in [1, ny + 3, 2 * ny + 4]:          # top three earth rows
        inv.strModel = m0.copy(); inv.strModel[cell] += h
        _, fp, _ = osamp.compDataGradient(mesh, data, inv, prior)
        inv.strModel = m0.copy(); inv.strModel[cell] -= h
        _, fm, _ = osamp.compDataGradient(mesh, data, inv, prior)
        fd = (fp - fm) / (2 * h)
        assert abs(fd - g[cell]) / np.abs(g).max() < 5e-3      # tiny 5-layer mesh: the BC approximations reach up


def test_rx_adjoint_closed_form_matches_sparse_L_Q():
    """the O(ny) reverse-mode recipe the CUDA kernel implements equals L^T d, Q^T d built
    from the reference's sparse algebra (dataFuncSens.jl)."""
    from tests.rx_closed_form import rx_adjoint_closed_form
    mesh, data, inv, prior = tiny_problem()
    pred, fwd = ofwd.MT2DFwdSolver(mesh, data)
    ny, nz = mesh.gridSize
    yNode = np.concatenate([[0.0], np.cumsum(mesh.yLen)]) - mesh.origin[0]
    zNode = np.concatenate([[0.0], np.cumsum(mesh.zLen)]) - mesh.origin[1]
    rxs = osens.preSetRxFieldSens(data.rxLoc, yNode, zNode, mesh.sigma)
    zid = rxs.zid
    rng = np.random.default_rng(11)
    d = rng.standard_normal(data.rxLoc.shape[0]) + 1j * rng.standard_normal(data.rxLoc.shape[0])
    om = 2 * np.pi * data.freqs[1]
    for mode, fld, fn in [(0, fwd.exTE, osens.getDataFuncSensTE), (1, fwd.hxTM, osens.getDataFuncSensTM)]:
        F01 = np.stack([fld[zid * (ny + 1):(zid + 1) * (ny + 1), 1], fld[(zid + 1) * (ny + 1):(zid + 2) * (ny + 1), 1]], 1)
        L, Q = fn(om, rxs, F01)
        s_ref = L.T @ d
        q_ref = Q.T @ d
        s0, s1, q = rx_adjoint_closed_form(mode, om, mesh.yLen, mesh.zLen[zid], rxs.sigma1, F01, data.rxLoc[:, 0], yNode, d)
        assert np.abs(s0 - s_ref[zid * (ny + 1):(zid + 1) * (ny + 1)]).max() / np.abs(s_ref).max() < 1e-12
        assert np.abs(s1 - s_ref[(zid + 1) * (ny + 1):(zid + 2) * (ny + 1)]).max() / np.abs(s_ref).max() < 1e-12
        assert np.abs(q - q_ref[zid * ny:(zid + 1) * ny]).max() / max(np.abs(q_ref).max(), 1e-300) < 1e-12


def test_oracle_reproduces_committed_fixture():
    """tests/golden/tiny_oracle.npz (tests/golden/make_oracle_fixture.py): the oracle must not drift."""
    import importlib.util
    import os
    from tests.helpers import GOLDEN
    spec = importlib.util.spec_from_file_location("make_oracle_fixture", os.path.join(GOLDEN, "make_oracle_fixture.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now = mod.compute()
    ref = np.load(os.path.join(GOLDEN, "tiny_oracle.npz"))
    for key in ("pred", "g", "m2", "p2"):
        scale = np.abs(ref[key]).max()
        assert np.abs(now[key] - ref[key]).max() <= 1e-12 * scale, key
    assert abs(now["phi"] - ref["phi"]) <= 1e-12 * abs(ref["phi"])
