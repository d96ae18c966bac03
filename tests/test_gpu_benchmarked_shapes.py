"""Oracle parity at the shapes bench.py times (cfg2: 200x100 cells — multifrontal solver (the default) and the register-window
band kernel T = 14 with two CTAs per system; cfg4: 800x300 cells, multifrontal solver) and a teacher-forced chain at 1e-9.  The oracle costs ~0.5 s per (frequency, mode) at
cfg2 and ~15 s at cfg4, so the number of frequencies is small; the kernels and launch configuration are the benchmarked ones.

Tolerances as in test_gpu_parity.py: 1e-9 wherever the reference's 1-D boundary recursion is well conditioned (few skin
depths: the low-frequency cases), max(1e-9, 20 x the oracle's own 1-ulp self-sensitivity) for a spectrum that includes 100 Hz."""
import copy

import numpy as np
import pytest

from tests.helpers import tiny_problem, to_oracle, to_product

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _oracle(mesh, data, inv, prior, m):
    from oracle import sampler as osamp
    om, od, oi, op = to_oracle(mesh, data, inv, prior)
    oi.strModel = m.copy()
    pred, phi, g = osamp.compDataGradient(om, od, oi, op)
    return (om, od, oi, op), pred, phi, g


def _self_sensitivity(octx, m, g0, pred0):
    from oracle import sampler as osamp
    om, od, oi, op = octx
    d2 = copy.copy(od)
    d2.freqs = np.nextafter(od.freqs, np.inf)
    oi.strModel = m.copy()
    pred1, _, g1 = osamp.compDataGradient(om, d2, oi, op)
    return np.abs(g1 - g0).max() / np.abs(g0).max(), (np.abs(pred1 - pred0) / np.abs(pred0)).max()


@pytest.mark.parametrize("split", [1, 0])
def test_cfg2_shape_full_spectrum(split, monkeypatch):
    """Round 1's benchmarked instantiation: band_factor_kernel<14> FM_OWN / FM_SEP + band_solve_kernel<14> SM_BACKZ_OWN
    (split = 1), and the unsplit kernel, on the stress model, frequencies 100 / 0.32 / 0.001 Hz."""
    from hmcmt2d_b200 import api, synthetic
    monkeypatch.setenv("HMCMT_SPLIT", str(split))
    monkeypatch.setenv("HMCMT_SOLVER", "band")
    mesh, data, inv, prior = synthetic.make_problem(200, 100, 3)
    m = synthetic.stress_model(inv)
    pl = api.Plan(mesh, data, inv, prior)
    assert pl.info(5) == 14 and pl.info(13) == split and pl.info(11) == 0
    pred, phi, g = pl.forward_gradient(m)
    assert pl.status() == 0
    pl.close()
    octx, opred, ophi, og = _oracle(mesh, data, inv, prior, m)
    sens_g, sens_p = _self_sensitivity(octx, m, og, opred)
    assert (np.abs(pred[0] - opred) / np.abs(opred)).max() < max(TOL, 20 * sens_p)
    assert abs(phi[0] - ophi) / abs(ophi) < max(TOL, 20 * sens_p)
    err = np.abs(g[0] - og) / np.abs(og).max()
    assert err.max() < max(TOL, 20 * sens_g), (err.max(), sens_g)


def test_cfg2_shape_default_solver_full_spectrum(monkeypatch):
    """The configuration bench.py times by default: multifrontal solver, stress model, frequencies 100 / 0.32 / 0.001 Hz."""
    from hmcmt2d_b200 import api, synthetic
    monkeypatch.delenv("HMCMT_SOLVER", raising=False)
    mesh, data, inv, prior = synthetic.make_problem(200, 100, 3)
    m = synthetic.stress_model(inv)
    pl = api.Plan(mesh, data, inv, prior)
    assert pl.info(11) == 1 and pl.info(5) == 0
    pred, phi, g = pl.forward_gradient(m)
    assert pl.status() == 0
    pl.close()
    octx, opred, ophi, og = _oracle(mesh, data, inv, prior, m)
    sens_g, sens_p = _self_sensitivity(octx, m, og, opred)
    assert (np.abs(pred[0] - opred) / np.abs(opred)).max() < max(TOL, 20 * sens_p)
    assert abs(phi[0] - ophi) / abs(ophi) < max(TOL, 20 * sens_p)
    err = np.abs(g[0] - og) / np.abs(og).max()
    assert err.max() < max(TOL, 20 * sens_g), (err.max(), sens_g)


@pytest.mark.parametrize("solver", ["band", "mf"])
def test_cfg2_shape_low_frequencies_hit_1e9(solver, monkeypatch):
    """Same mesh, 0.01 and 0.001 Hz: everything at 1e-9, for the band kernel (split) and the multifrontal solver."""
    from hmcmt2d_b200 import api, synthetic
    monkeypatch.setenv("HMCMT_SOLVER", solver)
    mesh, data, inv, prior = synthetic.make_problem(200, 100, 2, fmax_exp=-2.0, fmin_exp=-3.0)
    m = synthetic.stress_model(inv)
    pl = api.Plan(mesh, data, inv, prior)
    assert pl.info(11) == (1 if solver == "mf" else 0)
    pred, phi, g = pl.forward_gradient(m)
    pl.close()
    _, opred, ophi, og = _oracle(mesh, data, inv, prior, m)
    assert (np.abs(pred[0] - opred) / np.abs(opred)).max() < TOL
    assert abs(phi[0] - ophi) / abs(ophi) < TOL
    assert np.abs(g[0] - og).max() / np.abs(og).max() < TOL


def test_cfg4_shape_parity():
    """800x300 cells (N = 238 901 unknowns per system, multifrontal solver), one low frequency, TE + TM."""
    from hmcmt2d_b200 import api, synthetic
    mesh, data, inv, prior = synthetic.make_problem(800, 300, 1, fmax_exp=-2.5, fmin_exp=-2.5)
    m = synthetic.stress_model(inv)
    pl = api.Plan(mesh, data, inv, prior)
    assert pl.info(0) == 238901 and pl.info(11) == 1
    pred, phi, g = pl.forward_gradient(m)
    assert pl.status() == 0
    pl.close()
    octx, opred, ophi, og = _oracle(mesh, data, inv, prior, m)
    sens_g, sens_p = _self_sensitivity(octx, m, og, opred)
    assert (np.abs(pred[0] - opred) / np.abs(opred)).max() < max(TOL, 20 * sens_p)
    assert abs(phi[0] - ophi) / abs(ophi) < max(TOL, 20 * sens_p)
    assert np.abs(g[0] - og).max() / np.abs(og).max() < max(TOL, 20 * sens_g)


def test_teacher_forced_chain_1e9():
    """Accepted-sample chain with teacher forcing (SURVEY.md section 7 'Chain parity'): every trajectory starts from the
    ORACLE's current state, so chaos cannot accumulate and each proposal is held to 1e-9."""
    from hmcmt2d_b200 import api
    from oracle import sampler as osamp
    mesh, data, inv, prior = tiny_problem(seed=41)
    prior.dt = 0.02
    pm, pd, pi, pp = to_product(mesh, data, inv, prior)
    n = len(inv.strModel)
    st = osamp.make_streams(11, n, 3, prior.timestep)
    m_cur = inv.strModel.copy()
    p_cur = osamp.clip_momentum(st.z_init)
    for it in range(3):
        L = int(st.intsteps[it])
        om, op = osamp.proposeLeapfrog(m_cur.copy(), p_cur.copy(), mesh, data, inv, prior, L)
        gm, gp = api.proposeLeapfrog(api.HMCParameter(n, m_cur.copy(), p_cur.copy()), pm, pd, pi, pp, intstep=L)
        assert np.abs(gm - om).max() < TOL * max(1.0, np.abs(om).max()), it
        assert np.abs(gp - op).max() < TOL * max(1.0, np.abs(op).max()), it
        od, ok, oh, omn, opred = osamp.getHamiltonian(data, mesh, inv, prior, op)
        gd, gk, gh, gmn, gpred = api.getHamiltonian(pd, pm, pi, pp, api.HMCParameter(n, gm, gp))
        assert abs(gh - oh) < TOL * abs(oh) and (np.abs(gpred - opred) / np.abs(opred)).max() < TOL
        m_cur, p_cur = om.copy(), osamp.clip_momentum(st.z_momentum[it])      # teacher forcing: continue from the oracle's proposal


def test_jtvec_reuses_the_resident_factors():
    """hmcmt_jtvec after hmcmt_forward runs the adjoint only (compJacTMatVec.jl:220-224: the factors of the forward solve are
    reused): same numbers as the fused evaluation, at a fraction of its cost."""
    import time
    from hmcmt2d_b200 import api, synthetic
    mesh, data, inv, prior = synthetic.make_problem(200, 100, 30)
    m = synthetic.stress_model(inv)
    pl = api.Plan(mesh, data, inv, prior)
    pred, phi, g = pl.forward_gradient(m)
    v = (inv.dataW ** 2) * (pred[0] - inv.obsData)
    pl.forward(m=m, fields=False)
    gs = pl.jtvec(v)
    sigma_act = np.exp(m)
    assert np.abs(gs[0] * sigma_act - g[0]).max() / np.abs(g[0]).max() < 1e-12      # chain rule to log-sigma (HMCSampler.jl:306)
    t_full, t_jt = [], []
    for _ in range(3):
        t0 = time.perf_counter(); pl.forward_gradient(m); t_full.append(time.perf_counter() - t0)
        pl.forward(m=m, fields=False)
        t0 = time.perf_counter(); pl.jtvec(v); t_jt.append(time.perf_counter() - t0)
    assert min(t_jt) < 0.5 * min(t_full), (min(t_jt), min(t_full))
    pl.close()
