"""On-device leapfrog / Metropolis (K11) against the oracle's restatement of HMCSampler.jl with identical
injected random draws.  HMC is chaotic, so horizons are short ('Chain parity')."""
import copy

import numpy as np
import pytest

from tests.helpers import tiny_problem, to_product

pytestmark = pytest.mark.gpu


def test_leapfrog_trajectory_parity():
    from hmcmt2d_b200 import api
    from oracle import sampler as osamp
    mesh, data, inv, prior = tiny_problem(seed=21)
    pm, pd, pi, pp = to_product(mesh, data, inv, prior)
    rng = np.random.default_rng(22)
    m0 = inv.strModel + 0.2 * rng.standard_normal(len(inv.strModel))
    p0 = np.clip(rng.standard_normal(len(m0)), -2.5, 2.5)
    om, op = osamp.proposeLeapfrog(m0.copy(), p0.copy(), mesh, data, inv, prior, 3)
    par = api.HMCParameter(len(m0), m0.copy(), p0.copy())
    gm, gp = api.proposeLeapfrog(par, pm, pd, pi, pp, intstep=3)
    assert np.abs(gm - om).max() < 1e-9 * max(1.0, np.abs(om).max())
    assert np.abs(gp - op).max() < 1e-8 * max(1.0, np.abs(op).max())
    # Hamiltonian at the proposal (getHamiltonian HMCSampler.jl:358-397)
    od, ok, oh, omn, opred = osamp.getHamiltonian(data, mesh, inv, prior, op)
    gd, gk, gh, gmn, gpred = api.getHamiltonian(pd, pm, pi, pp, api.HMCParameter(len(m0), gm, gp))
    assert abs(gd - od) / od < 1e-8 and abs(gk - ok) / ok < 1e-8 and abs(gh - oh) / abs(oh) < 1e-8


def test_step_clip_and_bound_reflection():
    """Huge momentum: max|dt p| > 3 rescales the step (HMCSampler.jl:237-243) and models leaving
    [ln sigma_min, ln sigma_max] are reflected with a momentum flip (:515-559)."""
    from hmcmt2d_b200 import api
    from oracle import sampler as osamp
    mesh, data, inv, prior = tiny_problem(seed=23)
    prior.sigBounds = [2e-3, 5e-2]
    prior.dt = 0.5
    pm, pd, pi, pp = to_product(mesh, data, inv, prior)
    n = len(inv.strModel)
    rng = np.random.default_rng(24)
    m0 = np.clip(inv.strModel, np.log(2.1e-3), np.log(4.9e-2))
    p0 = 9.0 * rng.standard_normal(n)
    pl = api._plan_for(pm, pd, pi, pp)
    pl.set_state(m0, p0, m0)
    # one device step = drift (clip + reflect), gradient, full kick; compare against the oracle's pieces
    dm = prior.dt * p0
    assert np.abs(dm).max() > 3.0
    dm = dm / np.abs(dm).max() * 3.0
    om, op = osamp.checkParameterBound(m0 + dm, p0.copy(), prior)
    assert (op != p0).sum() > 0                                   # some parameters were reflected
    pl.leapfrog_steps_device(prior.dt, 1)
    gm, gp = pl.get_state()
    assert np.abs(gm[0] - om).max() < 1e-13
    assert np.all(gm[0] >= np.log(prior.sigBounds[0])) and np.all(gm[0] <= np.log(prior.sigBounds[1]))
    _, _, gdat = pl.forward_gradient(gm[0])
    gtot = gdat[0] + prior.regParam * (inv.Wm @ (gm[0] - m0))
    assert np.abs(gp[0] - (op - prior.dt * gtot)).max() < 1e-9 * np.abs(gp[0]).max()


@pytest.mark.parametrize("reuse", [False, True])
def test_short_chain_parity(reuse):
    """runHMCSampler with injected draws: models, statistics, accept flags, predicted data."""
    from hmcmt2d_b200 import api
    from oracle import sampler as osamp
    mesh, data, inv, prior = tiny_problem(seed=31)
    prior.dt = 0.02
    ns = 5
    n = len(inv.strModel)
    st = osamp.make_streams(7, n, ns, prior.timestep)
    pm, pd, pi, pp = to_product(mesh, data, inv, prior)
    omodel, ostats, odata = osamp.runHMCSampler(mesh, data, copy.copy(inv), prior, st, nsamples=ns)
    gs = api.RandomStreams(st.u_start, st.z_init, st.intsteps, st.u_accept, st.z_momentum)
    gmodel, gstat, gdata = api.runHMCSampler(pm, pd, pi, pp, gs, nsamples=ns, reuse_last_forward=reuse)
    assert gmodel.shape == omodel.shape == (n, ns) and gdata.shape == odata.shape
    assert np.array_equal(gstat.acceptstats, ostats["acceptstats"])
    assert gstat.nAccept == ostats["nAccept"] and gstat.nReject == ostats["nReject"]
    assert np.abs(gmodel - omodel).max() < 1e-7
    assert np.abs(gstat.hmstats - ostats["hmstats"]).max() / np.abs(ostats["hmstats"]).max() < 1e-7
    assert np.abs(gdata - odata).max() / np.abs(odata).max() < 1e-7


def test_nondiagonal_mass_matrix():
    """hmcprior.massType != "diagonal": M = Wm (setMassMatrix(invParam) HMCSampler.jl:478-489).  invM p (drift, kinetic energy) and
    sqrtM z (momentum draws) on the device against the oracle's dense Cholesky."""
    from hmcmt2d_b200 import api
    from oracle import sampler as osamp
    mesh, data, inv, prior = tiny_problem(seed=51)
    prior.massType = "nondiagonal"
    prior.dt = 0.02
    pm, pd, pi, pp = to_product(mesh, data, inv, prior)
    pp.massType = "nondiagonal"
    n = len(inv.strModel)
    invM, sqrtM = osamp.setMassMatrix(inv, prior)
    rng = np.random.default_rng(52)
    m0 = inv.strModel + 0.2 * rng.standard_normal(n)
    p0 = sqrtM @ np.clip(rng.standard_normal(n), -2.5, 2.5)
    om, op = osamp.proposeLeapfrog(m0.copy(), p0.copy(), mesh, data, inv, prior, 3, invM=invM)
    gm, gp = api.proposeLeapfrog(api.HMCParameter(n, m0.copy(), p0.copy()), pm, pd, pi, pp, intstep=3)
    assert np.abs(gm - om).max() < 1e-9 * max(1.0, np.abs(om).max())
    assert np.abs(gp - op).max() < 1e-8 * max(1.0, np.abs(op).max())
    od, ok, oh, omn, opred = osamp.getHamiltonian(data, mesh, inv, prior, op, invM=invM)
    gd, gk, gh, gmn, gpred = api.getHamiltonian(pd, pm, pi, pp, api.HMCParameter(n, gm, gp))
    assert abs(gk - ok) / ok < 1e-8 and abs(gh - oh) / abs(oh) < 1e-8
    # short chain: momentum draws through sqrtM, accept decisions, statistics
    mesh, data, inv, prior = tiny_problem(seed=53)
    prior.massType, prior.dt = "nondiagonal", 0.02
    pm, pd, pi, pp = to_product(mesh, data, inv, prior)
    pp.massType = "nondiagonal"
    ns = 4
    st = osamp.make_streams(9, n, ns, prior.timestep)
    omodel, ostats, odata = osamp.runHMCSampler(mesh, data, copy.copy(inv), prior, st, nsamples=ns)
    gs = api.RandomStreams(st.u_start, st.z_init, st.intsteps, st.u_accept, st.z_momentum)
    gmodel, gstat, gdata = api.runHMCSampler(pm, pd, pi, pp, gs, nsamples=ns, reuse_last_forward=False)
    assert np.array_equal(gstat.acceptstats, ostats["acceptstats"])
    assert np.abs(gmodel - omodel).max() < 1e-7
    assert np.abs(gstat.hmstats - ostats["hmstats"]).max() / np.abs(ostats["hmstats"]).max() < 1e-7
