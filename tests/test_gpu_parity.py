"""Parity of the CUDA hot path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances.  north_star: 1e-9 relative in complex FP64.  The reference's 1-D boundary solver
(mt1DAnalyticField / mt1DFieldSensMatrix) propagates up/down-going amplitudes top-down, which amplifies
rounding errors by exp(2 z/skin-depth): its own boundary values at depth change by up to ~1e-6 when an input
is perturbed by ONE ulp (measured below with the oracle itself).  Quantities that inherit this noise (deep
fields, the gradient on the deepest cell rows) are compared with tolerance max(1e-9, 20 x the oracle's own
1-ulp self-sensitivity); everything else is held to 1e-9."""
import copy

import numpy as np
import pytest

from tests.helpers import load_example, tiny_problem, to_product

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _oracle_grad(mesh, data, inv, prior, m):
    from oracle import sampler as osamp
    inv.strModel = m.copy()
    return osamp.compDataGradient(mesh, data, inv, prior)


def _self_sensitivity(mesh, data, inv, prior, m, g0, pred0):
    """Oracle vs oracle with the frequencies moved by one ulp."""
    d2 = copy.copy(data)
    d2.freqs = np.nextafter(data.freqs, np.inf)
    pred1, _, g1 = _oracle_grad(mesh, d2, inv, prior, m)
    return np.abs(g1 - g0).max() / np.abs(g0).max(), (np.abs(pred1 - pred0) / np.abs(pred0)).max()


def _gpu(mesh, data, inv, prior, m):
    from hmcmt2d_b200 import api
    pm, pd, pi, pp = to_product(mesh, data, inv, prior)
    pi.strModel = m.copy()
    pred, phi, g = api.compDataGradient(pm, pd, pi, pp)
    return pred, phi, g, (pm, pd, pi, pp)


def test_tiny_problem_full_parity():
    mesh, data, inv, prior = tiny_problem()
    m = inv.strModel.copy()
    opred, ophi, og = _oracle_grad(mesh, data, inv, prior, m)
    pred, phi, g, _ = _gpu(mesh, data, inv, prior, m)
    assert (np.abs(pred - opred) / np.abs(opred)).max() < TOL
    assert abs(phi - ophi) / abs(ophi) < TOL
    assert np.abs(g - og).max() / np.abs(og).max() < TOL


@pytest.mark.parametrize("solver", ["mf", "band"])
@pytest.mark.parametrize("name", ["dprism3d", "coprod2"])
def test_example_parity(name, solver, monkeypatch):
    """cfg1 (examples/dprism3d verbatim) and the cfg5 geometry (examples/coprod2), random log-normal model, on both solvers
    (multifrontal = the default at these sizes, register-window band kernel)."""
    monkeypatch.setenv("HMCMT_SOLVER", solver)
    mesh, data, inv, prior = load_example(name)
    rng = np.random.default_rng(1)
    m = np.log(0.01) + 0.7 * rng.standard_normal(len(inv.strModel))
    opred, ophi, og = _oracle_grad(mesh, data, inv, prior, m)
    sens_g, sens_p = _self_sensitivity(mesh, data, inv, prior, m, og, opred)
    pred, phi, g, _ = _gpu(mesh, data, inv, prior, m)
    assert (np.abs(pred - opred) / np.abs(opred)).max() < max(TOL, 20 * sens_p)
    assert abs(phi - ophi) / abs(ophi) < max(TOL, 20 * sens_p)
    err = np.abs(g - og) / np.abs(og).max()
    assert err.max() < max(TOL, 20 * sens_g), (err.max(), sens_g)
    # away from the noisy deep boundary (all but the 8 deepest cell rows) the gradient is held much tighter
    ny = mesh.gridSize[0]
    upper = np.arange(len(g)) < len(g) - 8 * ny
    assert err[upper].max() < max(TOL, 2 * sens_g)


def test_low_frequency_subset_hits_1e9():
    """When the mesh is only a few skin depths deep the 1-D recursions are well conditioned and the whole
    gradient meets 1e-9."""
    mesh, data, inv, prior = load_example("dprism3d")
    keep = data.freqID >= 9                        # 0.063, 0.025, 0.01 Hz
    d2 = copy.copy(data)
    d2.freqs = data.freqs[8:]
    d2.freqID = data.freqID[keep] - 8
    d2.rxID, d2.dtID = data.rxID[keep], data.dtID[keep]
    d2.dataID = np.ones(int(keep.sum()), bool)
    from oracle import sampler as osamp
    inv2 = osamp.setupInverseDataModel(mesh, [1e-8], inv.obsData[keep], 1.0 / inv.dataW[keep])
    rng = np.random.default_rng(2)
    m = np.log(0.01) + 0.5 * rng.standard_normal(len(inv2.strModel))
    opred, ophi, og = _oracle_grad(mesh, d2, inv2, prior, m)
    pred, phi, g, _ = _gpu(mesh, d2, inv2, prior, m)
    assert (np.abs(pred - opred) / np.abs(opred)).max() < TOL
    assert abs(phi - ophi) / abs(ophi) < TOL
    assert np.abs(g - og).max() / np.abs(og).max() < TOL


def test_system_export_pattern_is_bit_exact():
    """Sparsity pattern and DOF indexing must match bit-exactly (north_star)."""
    from oracle import forward as ofwd
    from oracle import operators as ops
    mesh, data, inv, prior = load_example("coprod2")
    m = inv.strModel + 0.3 * np.random.default_rng(4).standard_normal(len(inv.strModel))
    pred, phi, g, (pm, pd, pi, pp) = _gpu(mesh, data, inv, prior, m)
    from hmcmt2d_b200 import api
    pl = api._plan_for(pm, pd, pi, pp)
    mesh.sigma = inv.activeCell @ np.exp(m) + inv.bgModel
    ny, nz = mesh.gridSize
    ii, io = ops.getBoundaryIndex(ny, nz)
    for mode in (0, 1):
        coe = ofwd.assemble_mode(mesh, mode == 0, ii, io)
        for f in (0, 7):
            om = 2 * np.pi * data.freqs[f]
            A = (coe.rAii + 1j * om * coe.iAii).tocsc()
            A.sort_indices()
            colptr, rowval, nzval, rhs, bc = pl.export_system(mode, f)
            assert np.array_equal(colptr, A.indptr.astype(np.int64) + 1)          # 1-based CSC, as Julia hands it to MUMPS
            assert np.array_equal(rowval, A.indices.astype(np.int64) + 1)
            assert np.abs(nzval - A.data).max() / np.abs(A.data).max() < 1e-14
            assert np.abs(np.imag(nzval[rowval - 1 != np.repeat(np.arange(len(colptr) - 1), np.diff(colptr))])).max() == 0


def test_fields_and_jtvec_match_reference_interface():
    """MT2DFwdSolver -> (predData, MT2DFwdData) and compJacTMatVec with the reference's argument list."""
    from hmcmt2d_b200 import api
    from oracle import forward as ofwd
    from oracle import sensitivity as osens
    mesh, data, inv, prior = tiny_problem(seed=8)
    pm, pd, pi, pp = to_product(mesh, data, inv, prior)
    pred, fwd = api.MT2DFwdSolver(pm, pd)
    opred, ofw = ofwd.MT2DFwdSolver(mesh, data)
    assert pred.shape == opred.shape and fwd.exTE.shape == ofw.exTE.shape == ((mesh.gridSize[0] + 1) * (mesh.gridSize[1] + 1), 3)
    assert (np.abs(pred - opred) / np.abs(opred)).max() < TOL
    assert np.abs(fwd.exTE - ofw.exTE).max() / np.abs(ofw.exTE).max() < TOL
    assert np.abs(fwd.hxTM - ofw.hxTM).max() / np.abs(ofw.hxTM).max() < TOL
    rng = np.random.default_rng(9)
    v = rng.standard_normal(len(pred)) + 1j * rng.standard_normal(len(pred))
    g = api.compJacTMatVec(fwd.exTE, fwd.hxTM, v, pm, pd, pi.activeCell, fwd.AinvTE, fwd.AinvTM)
    og = osens.compJacTMatVec(ofw.exTE, ofw.hxTM, v, mesh, data, inv.activeCell, ofw.AinvTE, ofw.AinvTM)
    assert np.abs(g - og).max() / np.abs(og).max() < TOL
    # linearity in v (J^T is linear over the reals)
    v2 = rng.standard_normal(len(pred)) + 1j * rng.standard_normal(len(pred))
    g2 = api.compJacTMatVec(fwd.exTE, fwd.hxTM, v2, pm, pd, pi.activeCell, fwd.AinvTE, fwd.AinvTM)
    g12 = api.compJacTMatVec(fwd.exTE, fwd.hxTM, 2.0 * v - 0.5 * v2, pm, pd, pi.activeCell, fwd.AinvTE, fwd.AinvTM)
    assert np.abs(g12 - (2.0 * g - 0.5 * g2)).max() / np.abs(g).max() < 1e-12


def test_ragged_data_and_single_mode():
    """Missing data rows (dataID mask) and a TE-only survey."""
    from oracle import sampler as osamp
    mesh, data, inv, prior = tiny_problem(seed=5)
    rng = np.random.default_rng(6)
    keep = rng.random(len(inv.obsData)) > 0.3
    keep[:2] = True
    d2 = copy.copy(data)
    d2.freqID, d2.rxID, d2.dtID = data.freqID[keep], data.rxID[keep], data.dtID[keep]
    d2.dataID = keep.copy()
    inv2 = osamp.setupInverseDataModel(mesh, [1e-8], inv.obsData[keep], 1.0 / inv.dataW[keep])
    m = inv2.strModel.copy()
    opred, ophi, og = _oracle_grad(mesh, d2, inv2, prior, m)
    pred, phi, g, _ = _gpu(mesh, d2, inv2, prior, m)
    assert pred.shape == opred.shape == (int(keep.sum()),)
    assert (np.abs(pred - opred) / np.abs(opred)).max() < TOL and abs(phi - ophi) / ophi < TOL
    assert np.abs(g - og).max() / np.abs(og).max() < TOL
    te = data.dtID == 1
    d3 = copy.copy(data)
    d3.dataComp = ["ZXY"]
    d3.compTM = False
    d3.freqID, d3.rxID, d3.dtID = data.freqID[te], data.rxID[te], data.dtID[te]
    d3.dataID = np.ones(int(te.sum()), bool)
    inv3 = osamp.setupInverseDataModel(mesh, [1e-8], inv.obsData[te], 1.0 / inv.dataW[te])
    opred, ophi, og = _oracle_grad(mesh, d3, inv3, prior, m)
    pred, phi, g, _ = _gpu(mesh, d3, inv3, prior, m)
    assert (np.abs(pred - opred) / np.abs(opred)).max() < TOL and np.abs(g - og).max() / np.abs(og).max() < TOL


def test_batched_chains_equal_single_chains():
    from hmcmt2d_b200 import api
    mesh, data, inv, prior = tiny_problem(seed=12)
    pm, pd, pi, pp = to_product(mesh, data, inv, prior)
    rng = np.random.default_rng(13)
    ms = pi.strModel[None, :] + 0.3 * rng.standard_normal((3, len(pi.strModel)))
    pl3 = api.Plan(pm, pd, pi, pp, nChains=3)
    pred3, phi3, g3 = pl3.forward_gradient(ms)
    pl1 = api.Plan(pm, pd, pi, pp, nChains=1)
    for c in range(3):
        pred, phi, g = pl1.forward_gradient(ms[c])
        assert np.array_equal(pred[0], pred3[c]) and phi[0] == phi3[c] and np.array_equal(g[0], g3[c])      # bit-identical


def test_cfg5_finite_difference_check():
    """BASELINE.json configs[4]: examples/coprod2 geometry, forward / gradient finite-difference check on the GPU path.
    Central differences h = 1e-5 in ln(sigma) of the data misfit on probe cells away from the bottom 5 cell rows (the
    reference's boundary-derivative approximations live there) against the adjoint gradient."""
    from hmcmt2d_b200 import api
    mesh, data, inv, prior = load_example("coprod2")
    pm, pd, pi, pp = to_product(mesh, data, inv, prior)
    rng = np.random.default_rng(5)
    m = np.log(0.01) + 0.7 * rng.standard_normal(len(inv.strModel))
    pl = api.Plan(pm, pd, pi, pp)
    _, phi0, g = pl.forward_gradient(m)
    g = g[0]
    ny, nz = mesh.gridSize
    nair = len(mesh.airLayer)
    nrows = nz - nair                                    # earth cell rows (active cells are the earth cells, row-major)
    assert len(m) == ny * nrows
    rows = rng.integers(0, nrows - 5, size=10)
    cols = rng.integers(2, ny - 2, size=10)
    h = 1e-5
    worst = 0.0
    for r, c in zip(rows, cols):
        a = int(r) * ny + int(c)
        mp, mm = m.copy(), m.copy()
        mp[a] += h
        mm[a] -= h
        fd = (pl.forward_gradient(mp)[1][0] - pl.forward_gradient(mm)[1][0]) / (2 * h)
        worst = max(worst, abs(fd - g[a]) / np.abs(g).max())
    assert worst < 1e-4, worst
    pl.close()


def test_tiny_problem_against_committed_fixture():
    """The same seeded problem against the committed oracle vector (tests/golden/tiny_oracle.npz)."""
    import os
    from hmcmt2d_b200 import api
    from tests.helpers import GOLDEN
    ref = np.load(os.path.join(GOLDEN, "tiny_oracle.npz"))
    mesh, data, inv, prior = tiny_problem(seed=3)
    pm, pd, pi, pp = to_product(mesh, data, inv, prior)
    pi.strModel = ref["m"].copy()
    pred, phi, g = api.compDataGradient(pm, pd, pi, pp)
    assert (np.abs(pred - ref["pred"]) / np.abs(ref["pred"])).max() < TOL
    assert abs(phi - ref["phi"]) / abs(ref["phi"]) < TOL
    assert np.abs(g - ref["g"]).max() / np.abs(ref["g"]).max() < TOL
    m2, p2 = api.proposeLeapfrog(api.HMCParameter(len(ref["m"]), ref["m"].copy(), ref["p0"].copy()), pm, pd, pi, pp, intstep=2)
    assert np.abs(m2 - ref["m2"]).max() < TOL and np.abs(p2 - ref["p2"]).max() < TOL * max(1.0, np.abs(ref["p2"]).max())


def test_rho_phase_responses_forward_only():
    """DataType Rho_Pha: apparent resistivity and phase (compMTRespTE mt2DTE.jl:240-259, compMTRespTM mt2DTM.jl:224-242), interleaved
    and masked as MT2DFwdSolver.jl:191-205 does.  Forward only — the reference's sensitivity code tests "Rho_Phs" and never
    reaches that data type (compJacTMatVec.jl:104), so the gradient refuses it."""
    from hmcmt2d_b200 import api, fileio
    from oracle import fileio as ofio
    from oracle import forward as ofwd
    mesh, data, inv, prior = tiny_problem(seed=17)
    nF, nRx = len(data.freqs), data.rxLoc.shape[0]
    comps = ["RhoXY", "PhsXY", "RhoYX", "PhsYX"]
    f, r, c = np.meshgrid(np.arange(1, nF + 1), np.arange(1, nRx + 1), np.arange(1, 5), indexing="ij")
    keep = np.random.default_rng(3).random(nF * nRx * 4) > 0.25                      # ragged: some rows absent
    od = ofio.MTData(data.rxLoc, data.freqs, "Rho_Pha", comps, r.ravel()[keep].astype(np.int64), f.ravel()[keep].astype(np.int64),
                     c.ravel()[keep].astype(np.int64), keep.copy(), True, True)
    opred, _ = ofwd.MT2DFwdSolver(mesh, od)
    pm, _, _, _ = to_product(mesh, data, inv, prior)
    pd = fileio.MTData(np.array(od.rxLoc), np.array(od.freqs), "Rho_Pha", comps, np.array(od.rxID), np.array(od.freqID),
                       np.array(od.dtID), np.array(od.dataID), True, True)
    pred, fwd = api.MT2DFwdSolver(pm, pd)
    assert pred.dtype == np.float64 and pred.shape == opred.shape == (int(keep.sum()),)
    assert np.abs(pred - opred).max() / np.abs(opred).max() < TOL
    assert (np.abs(pred - opred) / np.maximum(np.abs(opred), 1e-30)).max() < 1e-8           # phases near 0 / 180 included
    with pytest.raises(NotImplementedError):
        api.compJacTMatVec(fwd.exTE, fwd.hxTM, np.zeros(len(pred)), pm, pd, None, fwd.AinvTE, fwd.AinvTM)
    # TE-only survey: [rho, phs] pairs
    te = np.isin(od.dtID, (1, 2))
    od1 = ofio.MTData(data.rxLoc, data.freqs, "Rho_Pha", comps[:2], od.rxID[te], od.freqID[te], od.dtID[te],
                      keep.reshape(nF, nRx, 4)[:, :, :2].reshape(-1).copy(), True, False)
    opred1, _ = ofwd.MT2DFwdSolver(mesh, od1)
    pd1 = fileio.MTData(np.array(od1.rxLoc), np.array(od1.freqs), "Rho_Pha", comps[:2], np.array(od1.rxID), np.array(od1.freqID),
                        np.array(od1.dtID), np.array(od1.dataID), True, False)
    pred1, _ = api.MT2DFwdSolver(pm, pd1)
    assert np.abs(pred1 - opred1).max() / np.abs(opred1).max() < TOL


def test_stale_factor_handle_is_refused():
    """One set of factors is resident per plan: the MT2DFwdData of an earlier evaluation must not silently return the
    gradient at a later model (the reference's MT2DFwdData owns its factors, MT2DFwdSolver.jl:44-53)."""
    from hmcmt2d_b200 import api
    mesh, data, inv, prior = tiny_problem(seed=8)
    pm, pd, pi, pp = to_product(mesh, data, inv, prior)
    pred_a, fwd_a = api.MT2DFwdSolver(pm, pd)
    sig_a = pm.sigma.copy()
    v = np.ones(len(pred_a), dtype=complex)
    g_a = api.compJacTMatVec(fwd_a.exTE, fwd_a.hxTM, v, pm, pd, pi.activeCell, fwd_a.AinvTE, fwd_a.AinvTM)
    pm.sigma = sig_a * 1.3
    pred_b, fwd_b = api.MT2DFwdSolver(pm, pd)
    with pytest.raises(RuntimeError):
        api.compJacTMatVec(fwd_a.exTE, fwd_a.hxTM, v, pm, pd, pi.activeCell, fwd_a.AinvTE, fwd_a.AinvTM)
    g_b = api.compJacTMatVec(fwd_b.exTE, fwd_b.hxTM, v, pm, pd, pi.activeCell, fwd_b.AinvTE, fwd_b.AinvTM)
    assert np.abs(g_b - g_a).max() > 1e-6 * np.abs(g_a).max()
    # a selector with the same number of columns but other cells is mapped by cell index, not by shape
    import scipy.sparse as sp
    nAC = pi.activeCell.shape[1]
    perm = np.random.default_rng(0).permutation(nAC)
    P2 = sp.csr_matrix(pi.activeCell.tocsc()[:, perm])
    g_p = api.compJacTMatVec(fwd_b.exTE, fwd_b.hxTM, v, pm, pd, P2, fwd_b.AinvTE, fwd_b.AinvTM)
    assert np.array_equal(g_p, g_b[perm])


def test_explicit_jacobian_matches_reference_derivation():
    """compJacMat (compJacMat.jl:7-381): the explicit complex Jacobian against the oracle's restatement, and its consistency with
    J^T v (compJacTMatVec) on the device."""
    from hmcmt2d_b200 import api
    from oracle import forward as ofwd
    from oracle import jacobian as ojac
    mesh, data, inv, prior = tiny_problem(seed=19)
    pm, pd, pi, pp = to_product(mesh, data, inv, prior)
    pred, fwd = api.MT2DFwdSolver(pm, pd)
    J = api.compJacMat(fwd.exTE, fwd.hxTM, pm, pd, pi.activeCell, fwd.AinvTE, fwd.AinvTM)
    opred, ofw = ofwd.MT2DFwdSolver(mesh, data)
    oJ = ojac.compJacMat(ofw.exTE, ofw.hxTM, mesh, data, inv.activeCell, ofw.AinvTE, ofw.AinvTM)
    assert J.shape == oJ.shape == (len(pred), inv.activeCell.shape[1])
    assert np.abs(J - oJ).max() / np.abs(oJ).max() < TOL
    rng = np.random.default_rng(2)
    v = rng.standard_normal(len(pred)) + 1j * rng.standard_normal(len(pred))
    g = api.compJacTMatVec(fwd.exTE, fwd.hxTM, v, pm, pd, pi.activeCell, fwd.AinvTE, fwd.AinvTM)
    assert np.abs(g - np.real(J.T @ np.conj(v))).max() / np.abs(g).max() < 1e-11
    JT = api.compJacTMat(fwd.exTE, fwd.hxTM, pm, pd, pi.activeCell, fwd.AinvTE, fwd.AinvTM)
    assert JT.shape == J.shape[::-1] and np.array_equal(JT, J.T)
    # the responses of the plan are intact after the adjoint passes
    pred2, _ = api.MT2DFwdSolver(pm, pd)
    assert np.array_equal(pred2, pred)


def test_total_gradient_entry_point():
    """hmcmt_forward_gradient_total = data gradient + beta Wm (m - m_ref) (the sum proposeLeapfrog forms, HMCSampler.jl:240-262)."""
    from hmcmt2d_b200 import api
    from tests.helpers import tiny_problem, to_product
    mesh, data, inv, prior = tiny_problem(seed=5)
    pm, pd, pi, pp = to_product(mesh, data, inv, prior)
    pl = api.Plan(pm, pd, pi, pp)
    rng = np.random.default_rng(1)
    mref = pi.strModel.copy()
    m = mref + 0.3 * rng.standard_normal(len(mref))
    pl.set_state(m, np.zeros_like(m), mref)
    pred, phi, g = pl.forward_gradient(m)
    pred2, phi2, gt = pl.forward_gradient(m, total=True)
    assert np.array_equal(pred, pred2) and np.array_equal(phi, phi2)
    want = g[0] + pp.regParam * (pi.Wm @ (m - mref))
    assert np.abs(gt[0] - want).max() <= 1e-12 * np.abs(want).max()
    pl.close()
