"""Nested-dissection multifrontal solver (mf_symbolic.h / mf_kernels.cuh) on the GPU, and the frequency-sharded step.
Solver boundary: the reference's own criterion (MUMPS/test/testDivGrad.jl: relative residual < 1e-14) on matrices the
register-window band kernel cannot take.  Full path: meshes wider than 104 unknowns per line against the CPU oracle at 1e-9
(north_star)."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def relres(A, x, b):
    return np.linalg.norm(A @ x - b) / np.linalg.norm(b)


def _stencil(nl, nf, rng):
    N = nl * nf
    d = 4 + rng.random(N) + 1j * rng.random(N)
    e1, e2 = -rng.random(N), -rng.random(N)
    e1[np.arange(N) % nf == 0] = 0
    return sp.diags([d, e1[1:], e1[1:], e2[nf:], e2[nf:]], [0, -1, 1, -nf, nf], format="csc")


def test_wide_stencil_systems():
    from hmcmt2d_b200 import lib
    rng = np.random.default_rng(2)
    for nl, nf in [(4, 105), (7, 130), (5, 200), (4, 299), (3, 320), (40, 500), (300, 130)]:
        A = _stencil(nl, nf, rng)
        rhs = rng.standard_normal(A.shape[0]) + 1j * rng.standard_normal(A.shape[0])
        assert relres(A, lib.solveMUMPS(A, rhs, 1), rhs) < 1e-14, (nl, nf)


@pytest.mark.parametrize("fsmall,leaf", [(144, 16), (0, 16), (48, 8), (144, 1), (80, 64)])
def test_front_size_classes(fsmall, leaf, monkeypatch):
    """Every kernel family: all fronts in shared memory, all fronts through the global-memory path (assembly, diagonal-block
    inversion, DMMA GEMMs, chunked pivots), mixed; single-node and 64-node leaves.  Forced onto matrices the band kernel would
    otherwise take."""
    from hmcmt2d_b200 import lib
    monkeypatch.setenv("HMCMT_SHIM_SOLVER", "mf")
    monkeypatch.setenv("HMCMT_MF_FSMALL", str(fsmall))
    monkeypatch.setenv("HMCMT_MF_LEAF", str(leaf))
    rng = np.random.default_rng(5)
    for nl, nf in [(1, 1), (5, 3), (3, 8), (20, 13), (64, 64), (130, 97)]:
        A = _stencil(nl, nf, rng) if nl * nf > 1 else sp.csc_matrix(np.array([[2.0 + 1.0j]]))
        rhs = rng.standard_normal((A.shape[0], 3)) + 1j * rng.standard_normal((A.shape[0], 3))
        x = lib.solveMUMPS(A, rhs, 2)
        assert max(relres(A, x[:, i], rhs[:, i]) for i in range(3)) < 1e-14, (nl, nf)


def test_full_band_matrix_and_ragged_size():
    """Dense band (every diagonal populated: fronts far denser than a grid's), ragged N, several right-hand sides, real twin."""
    from hmcmt2d_b200 import lib
    rng = np.random.default_rng(3)
    N, b = 1003, 120
    offs = list(range(1, b + 1))
    lower = [rng.standard_normal(N - o) + 1j * rng.standard_normal(N - o) for o in offs]
    A = sp.diags([8 * b * (1 + rng.random(N)) + 1j * rng.random(N)] + lower + lower, [0] + [-o for o in offs] + offs, format="csc")
    rhs = rng.standard_normal((N, 3)) + 1j * rng.standard_normal((N, 3))
    F = lib.factorMUMPS(A, 2)
    x = lib.applyMUMPS(F, rhs)
    assert max(relres(A, x[:, i], rhs[:, i]) for i in range(3)) < 1e-14
    lib.destroyMUMPS(F)
    Ar = A.real.tocsc()
    xr = lib.solveMUMPS(Ar, rhs[:, 0].real.copy(), 1)
    assert xr.dtype == np.float64 and relres(Ar, xr, rhs[:, 0].real) < 1e-14


def test_disconnected_and_arrow_patterns():
    """Graphs the level-set bisection must survive: two disconnected grids in one matrix, and an arrow matrix (one dense row)."""
    from hmcmt2d_b200 import lib
    rng = np.random.default_rng(6)
    A = sp.block_diag([_stencil(30, 20, rng), _stencil(12, 45, rng)], format="csc")
    rhs = rng.standard_normal(A.shape[0]) + 1j * rng.standard_normal(A.shape[0])
    import os
    os.environ["HMCMT_SHIM_SOLVER"] = "mf"
    try:
        assert relres(A, lib.solveMUMPS(A, rhs, 1), rhs) < 1e-14
        n = 400
        B = sp.lil_matrix((n, n), dtype=complex)
        B.setdiag(4.0 + rng.random(n) + 1j * rng.random(n))
        B[0, 0] = 400.0
        B[n - 1, :n - 1] = -rng.random(n - 1) * 0.01
        B[:n - 1, n - 1] = B[n - 1, :n - 1].T
        B[n - 1, n - 1] = 8.0
        B = B.tocsc()
        rhs = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        assert relres(B, lib.solveMUMPS(B, rhs, 2), rhs) < 1e-14
    finally:
        os.environ.pop("HMCMT_SHIM_SOLVER", None)


def test_wide_mesh_parity_with_oracle():
    """ny-1 = 123, nz-1 = 117 > 104: the plan takes the multifrontal solver."""
    from hmcmt2d_b200 import api, synthetic
    from oracle import sampler as osamp
    from tests.helpers import to_oracle
    mesh, data, inv, prior = synthetic.make_problem(124, 118, 2, nRx=10)
    m = synthetic.stress_model(inv)
    pl = api.Plan(mesh, data, inv, prior)
    assert pl.info(4) == 117 and pl.info(5) == 0 and pl.info(11) == 1
    pred, phi, g = pl.forward_gradient(m)
    om, od, oi, op = to_oracle(mesh, data, inv, prior)
    oi.strModel = m.copy()
    opred, ophi, og = osamp.compDataGradient(om, od, oi, op)
    assert (np.abs(pred[0] - opred) / np.abs(opred)).max() < 1e-9
    assert abs(phi[0] - ophi) / abs(ophi) < 1e-9
    assert np.abs(g[0] - og).max() / np.abs(og).max() < 1e-9
    pl.close()


def test_multifrontal_equals_band_kernel(monkeypatch):
    """The two solvers on the same narrow mesh: identical responses / gradient to round-off, batched chains included."""
    from hmcmt2d_b200 import api, synthetic
    mesh, data, inv, prior = synthetic.make_problem(60, 40, 3, nRx=8)
    ms = np.stack([synthetic.stress_model(inv, seed=s) for s in (1, 2)])
    monkeypatch.setenv("HMCMT_SOLVER", "band")
    pb = api.Plan(mesh, data, inv, prior, nChains=2)
    assert pb.info(11) == 0
    pred0, phi0, g0 = pb.forward_gradient(ms)
    pb.close()
    monkeypatch.setenv("HMCMT_SOLVER", "mf")
    pm = api.Plan(mesh, data, inv, prior, nChains=2)
    assert pm.info(11) == 1
    pred1, phi1, g1 = pm.forward_gradient(ms)
    pm.close()
    assert (np.abs(pred1 - pred0) / np.abs(pred0)).max() < 1e-11
    assert np.abs(phi1 - phi0).max() / np.abs(phi0).max() < 1e-10
    assert np.abs(g1 - g0).max() / np.abs(g0).max() < 1e-10


def test_execution_modes_are_bit_identical(monkeypatch):
    """How an evaluation is issued must not change a single bit of it: one stream / three groups of systems on their own streams,
    eager launches / the replayed CUDA graph (from the second evaluation of a plan on), pruned / complete forward eliminations.
    Device-resident leapfrog steps (several evaluations per plan, so the graph is captured and replayed) and the host-buffer
    entry point."""
    from hmcmt2d_b200 import api, synthetic
    mesh, data, inv, prior = synthetic.make_problem(70, 50, 4, nRx=9)
    m0 = synthetic.stress_model(inv)
    p0 = np.clip(np.random.default_rng(0).standard_normal(len(m0)), -2.5, 2.5)
    results = []
    for groups, graph, prune in [("1", "0", "0"), ("3", "0", "1"), ("1", "1", "1"), ("3", "1", "1"), ("8", "1", "0")]:
        monkeypatch.setenv("HMCMT_SOLVER", "mf")
        monkeypatch.setenv("HMCMT_GROUPS", groups)
        monkeypatch.setenv("HMCMT_GRAPH", graph)
        monkeypatch.setenv("HMCMT_MF_PRUNE", prune)
        pl = api.Plan(mesh, data, inv, prior)
        assert pl.info(11) == 1
        pl.set_state(m0, p0, m0)
        pl.leapfrog_steps_device(prior.dt, 4)
        assert pl.status() == 0
        m, p = pl.get_state()
        outs = [pl.forward_gradient(m0) for _ in range(3)]          # eager, captured, replayed
        for o in outs[1:]:
            assert all(np.array_equal(a, b) for a, b in zip(o, outs[0]))
        results.append((m.copy(), p.copy()) + tuple(np.array(a) for a in outs[0]))
        pl.close()
    for r in results[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(r, results[0]))


def test_error_flags_through_graph_replay(monkeypatch):
    """A non-finite model makes some pivot block singular: the call that caused it returns -10 (eager launch or graph replay), the
    flags are cleared, and the next valid evaluation of the same plan is correct again."""
    from hmcmt2d_b200 import api, lib, synthetic
    monkeypatch.setenv("HMCMT_SOLVER", "mf")
    mesh, data, inv, prior = synthetic.make_problem(64, 45, 2, nRx=8)
    m = synthetic.stress_model(inv)
    pl = api.Plan(mesh, data, inv, prior)
    good = [pl.forward_gradient(m) for _ in range(3)]          # eager, capture, replay
    bad = m.copy()
    bad[: len(bad) // 2] = np.nan
    with pytest.raises(lib.HmcmtError) as e:
        pl.forward_gradient(bad)
    assert e.value.code == -10
    again = pl.forward_gradient(m)
    assert all(np.array_equal(a, b) for a, b in zip(again, good[0]))
    pl.close()


@pytest.mark.parametrize("leaf,cross,push", [(16, 0, 0), (16, 0, 2), (9, 0, 7), (16, 13, 2), (36, 26, 0), (1, 0, 0)])
def test_ordering_knobs_against_oracle(leaf, cross, push, monkeypatch):
    """Leaf boxes, cross-shaped separators and the tile alignment of the leaves (excess unknowns handed to the separator above)
    only change the elimination order: every variant must reproduce the oracle to 1e-9."""
    from hmcmt2d_b200 import api, synthetic
    from oracle import sampler as osamp
    from tests.helpers import to_oracle
    monkeypatch.setenv("HMCMT_SOLVER", "mf")
    monkeypatch.setenv("HMCMT_MF_LEAF", str(leaf))
    monkeypatch.setenv("HMCMT_MF_CROSS", str(cross))
    monkeypatch.setenv("HMCMT_MF_PUSH", str(push))
    mesh, data, inv, prior = synthetic.make_problem(64, 45, 2, nRx=8)
    m = synthetic.stress_model(inv)
    pl = api.Plan(mesh, data, inv, prior)
    pred, phi, g = pl.forward_gradient(m)
    pl.close()
    om, od, oi, op = to_oracle(mesh, data, inv, prior)
    oi.strModel = m.copy()
    opred, ophi, og = osamp.compDataGradient(om, od, oi, op)
    assert (np.abs(pred[0] - opred) / np.abs(opred)).max() < 1e-9
    assert abs(phi[0] - ophi) / abs(ophi) < 1e-9
    assert np.abs(g[0] - og).max() / np.abs(og).max() < 1e-9


def test_frequency_sharded_steps_match_unsharded():
    """Two ranks' plans in one process, the all-reduce done by hand on the exchange buffers: the sharded leapfrog steps
    (partial -> sum -> finish) reproduce the unsharded device loop."""
    import torch
    from hmcmt2d_b200 import api, synthetic
    mesh, data, inv, prior = synthetic.make_problem(40, 30, 5, nRx=8)
    m0 = synthetic.stress_model(inv)
    p0 = np.clip(np.random.default_rng(0).standard_normal(len(m0)), -2.5, 2.5)
    full = api.Plan(mesh, data, inv, prior)
    shards = [api.FreqShardedPlan(mesh, data, inv, prior, r, 2) for r in range(2)]
    views = [s._exchange_tensor() for s in shards]

    def sharded_steps(n):
        for _ in range(n):
            for s in shards:
                s.plan.step_partial(prior.dt)
                s.sync()
            tot = views[0] + views[1]
            for v in views:
                v.copy_(tot)
            torch.cuda.synchronize()
            for s in shards:
                s.plan.step_finish(prior.dt)

    # one step: only the summation order over frequencies differs -> round-off agreement.  Three steps: the states differ
    # by ulps after the first one and the reference's 1-D boundary recursion amplifies that (up to ~1e-7
    # of max|g| per ulp), so later steps are held to the noise-bounded tolerance of test_gpu_parity.py.
    for nsteps, tol in ((1, 1e-11), (3, 1e-6)):
        full.set_state(m0, p0, m0)
        full.leapfrog_steps_device(prior.dt, nsteps)
        mf, pf = full.get_state()
        for s in shards:
            s.set_state(m0, p0, m0)
        sharded_steps(nsteps)
        for s in shards:
            ms, ps = s.get_state()
            assert np.abs(ms - mf).max() < tol and np.abs(ps - pf).max() < tol * max(1.0, np.abs(pf).max()), nsteps
    # host-buffer evaluation of one shard + its complement = the full evaluation
    pred, phi, g = full.forward_gradient(m0)
    parts = [s.plan.forward_gradient(m0) for s in shards]
    assert np.abs(parts[0][2] + parts[1][2] - g).max() < 1e-9 * np.abs(g).max()
    assert abs(parts[0][1][0] + parts[1][1][0] - phi[0]) < 1e-9 * abs(phi[0])
    merged = np.zeros(len(inv.obsData), complex)
    for s, part in zip(shards, parts):
        merged[s.rows] = part[0][0]
    assert (np.abs(merged - pred[0]) / np.abs(pred[0])).max() < 1e-12
