"""GPU bring-up of the multifrontal solver (prints diagnostics, asserts nothing): shim residuals for small / large-front
configurations, then the plan path against the band path and the oracle.   python tools/dev/t_mf.py [stage ...]"""
import os
import sys
import time
import traceback

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hmcmt2d_b200 import lib  # noqa: E402


def relres(A, x, b):
    if b.ndim == 1:
        return np.linalg.norm(A @ x - b) / np.linalg.norm(b)
    return max(np.linalg.norm(A @ x[:, i] - b[:, i]) / np.linalg.norm(b[:, i]) for i in range(b.shape[1]))


def stencil(nl, nf, rng, yfast=False):
    N = nl * nf
    d = 4 + rng.random(N) + 1j * rng.random(N)
    e1, e2 = -rng.random(N), -rng.random(N)
    e1[np.arange(N) % nf == 0] = 0
    return sp.diags([d, e1[1:], e1[1:], e2[nf:], e2[nf:]], [0, -1, 1, -nf, nf], format="csc")


def ddx(n):
    return sp.diags([-np.ones(n), np.ones(n)], [0, 1], shape=(n, n + 1))


def divgrad(n1, n2, n3):
    I = sp.identity
    Div = sp.hstack([sp.kron(I(n3), sp.kron(I(n2), ddx(n1))), sp.kron(I(n3), sp.kron(ddx(n2), I(n1))), sp.kron(ddx(n3), sp.kron(I(n2), I(n1)))])
    return (Div @ Div.T).tocsc()


def stage_shim():
    rng = np.random.default_rng(0)
    os.environ["HMCMT_SHIM_SOLVER"] = "mf"
    for fsmall, leaf, shapes in [(144, 16, [(5, 3), (3, 8), (20, 13), (40, 25), (64, 64)]), (0, 16, [(5, 3), (20, 13), (40, 25), (64, 64)]),
                                 (144, 16, [(199, 99), (99, 199), (300, 130)]), (48, 8, [(64, 64)])]:
        os.environ["HMCMT_MF_FSMALL"], os.environ["HMCMT_MF_LEAF"] = str(fsmall), str(leaf)
        for nl, nf in shapes:
            try:
                A = stencil(nl, nf, rng)
                b = rng.standard_normal(A.shape[0]) + 1j * rng.standard_normal(A.shape[0])
                t0 = time.time()
                F = lib.factorMUMPS(A, 1)
                t1 = time.time()
                x = lib.applyMUMPS(F, b)
                t2 = time.time()
                F2 = lib.factorMUMPS(A, 1)          # cached symbolic
                t3 = time.time()
                lib.destroyMUMPS(F); lib.destroyMUMPS(F2)
                print(f"[shim] fsmall {fsmall:3d} leaf {leaf:2d} grid {nl}x{nf}: resid {relres(A, x, b):.2e}  factor {t1 - t0:.3f}s (2nd {t3 - t2:.3f}s) solve {t2 - t1:.3f}s", flush=True)
            except Exception as e:
                print(f"[shim] fsmall {fsmall} leaf {leaf} grid {nl}x{nf}: FAILED {e!r}", flush=True)
    os.environ["HMCMT_MF_FSMALL"], os.environ["HMCMT_MF_LEAF"] = "144", "16"
    for dims in [(10, 10, 16), (32, 32, 16)]:
        try:
            A = divgrad(*dims)
            n = A.shape[0]
            b = rng.standard_normal((n, 10))
            t0 = time.time()
            x = lib.solveMUMPS(A, b, 1)
            t1 = time.time()
            Ac = (A + 1j * sp.diags(rng.random(n))).tocsc()
            bc = rng.standard_normal((n, 10)) + 1j * rng.standard_normal((n, 10))
            xc = lib.solveMUMPS(Ac, bc, 2)
            print(f"[shim] divgrad {dims}: real resid {relres(A, x, b):.2e} ({t1 - t0:.2f}s)  complex resid {relres(Ac, xc, bc):.2e} ({time.time() - t1:.2f}s)", flush=True)
        except Exception as e:
            print(f"[shim] divgrad {dims}: FAILED {e!r}", flush=True)
    os.environ.pop("HMCMT_SHIM_SOLVER")


def stage_one():
    rng = np.random.default_rng(0)
    os.environ["HMCMT_SHIM_SOLVER"] = "mf"
    for nl, nf in [(5, 3), (20, 13)]:
        A = stencil(nl, nf, rng)
        b = rng.standard_normal(A.shape[0]) + 1j * rng.standard_normal(A.shape[0])
        x = lib.solveMUMPS(A, b, 1)
        print(f"[one] grid {nl}x{nf}: resid {relres(A, x, b):.2e}", flush=True)
    os.environ["HMCMT_MF_FSMALL"] = "0"
    for nl, nf in [(6, 3), (21, 13)]:
        A = stencil(nl, nf, rng)
        b = rng.standard_normal(A.shape[0]) + 1j * rng.standard_normal(A.shape[0])
        x = lib.solveMUMPS(A, b, 1)
        print(f"[one] big path grid {nl}x{nf}: resid {relres(A, x, b):.2e}", flush=True)


def plan_eval(ny, nz, nf, solver, nrx=10, reps=3):
    from hmcmt2d_b200 import api, synthetic
    if solver:
        os.environ["HMCMT_SOLVER"] = solver
    else:
        os.environ.pop("HMCMT_SOLVER", None)
    mesh, data, inv, prior = synthetic.make_problem(ny, nz, nf, nRx=nrx)
    m = synthetic.stress_model(inv)
    t0 = time.time()
    pl = api.Plan(mesh, data, inv, prior)
    t1 = time.time()
    out = pl.forward_gradient(m)
    ts = []
    for _ in range(reps):
        t2 = time.time()
        out = pl.forward_gradient(m)
        ts.append(time.time() - t2)
    info = dict(T=pl.info(5), mf=pl.info(11), factor_MB=pl.info(9) / 1e6, flops=pl.info(12), launches=pl.info(10))
    pl.close()
    os.environ.pop("HMCMT_SOLVER", None)
    return out, (mesh, data, inv, prior, m), dict(plan_s=t1 - t0, eval_s=min(ts), **info)


def cmp(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def stage_plan_small():
    for ny, nz, nf in [(30, 24, 2), (60, 40, 3), (124, 118, 2)]:
        try:
            (p1, f1, g1), ctx, i1 = plan_eval(ny, nz, nf, "mf")
            print(f"[plan] {ny}x{nz} nf {nf} mf: {i1}", flush=True)
            if min(ny, nz) - 1 <= 104:
                (p0, f0, g0), _, i0 = plan_eval(ny, nz, nf, None)
                print(f"[plan] {ny}x{nz} band: {i0}")
                print(f"[plan] {ny}x{nz} mf vs band: pred {cmp(p1, p0):.2e} phi {abs(f1[0] - f0[0]) / abs(f0[0]):.2e} grad {cmp(g1, g0):.2e}", flush=True)
            from oracle import sampler as osamp
            from tests.helpers import to_oracle
            mesh, data, inv, prior, m = ctx
            om, od, oi, op = to_oracle(mesh, data, inv, prior)
            oi.strModel = m.copy()
            opred, ophi, og = osamp.compDataGradient(om, od, oi, op)
            print(f"[plan] {ny}x{nz} mf vs oracle: pred {float((np.abs(p1[0] - opred) / np.abs(opred)).max()):.2e} phi {abs(f1[0] - ophi) / abs(ophi):.2e} grad {cmp(g1[0], og):.2e}", flush=True)
        except Exception:
            traceback.print_exc()


def stage_cfg2():
    try:
        (p0, f0, g0), _, i0 = plan_eval(200, 100, 30, None, nrx=40)
        print(f"[cfg2] band: {i0}", flush=True)
        for leaf in (16, 8, 32):
            os.environ["HMCMT_MF_LEAF"] = str(leaf)
            (p1, f1, g1), _, i1 = plan_eval(200, 100, 30, "mf", nrx=40)
            print(f"[cfg2] mf leaf {leaf}: {i1}")
            print(f"[cfg2] mf vs band: pred {cmp(p1, p0):.2e} phi {abs(f1[0] - f0[0]) / abs(f0[0]):.2e} grad {cmp(g1, g0):.2e}", flush=True)
        os.environ.pop("HMCMT_MF_LEAF", None)
    except Exception:
        traceback.print_exc()


def stage_cfg4():
    try:
        for nf in (2, 8):
            (p1, f1, g1), _, i1 = plan_eval(800, 300, nf, None, nrx=40, reps=2)
            print(f"[cfg4] nfreq {nf}: {i1}", flush=True)
    except Exception:
        traceback.print_exc()


def stage_sweep():
    for leaf in (8, 16, 24, 32, 48, 64):
        os.environ["HMCMT_MF_LEAF"] = str(leaf)
        try:
            _, _, i4 = plan_eval(800, 300, 8, None, nrx=40, reps=3)
            _, _, i2 = plan_eval(200, 100, 30, "mf", nrx=40, reps=5)
            print(f"[sweep] leaf {leaf}: cfg4x16 {i4['eval_s'] * 1e3:.2f} ms (factor {i4['factor_MB']:.0f} MB, {i4['flops']:.3e} flop)   "
                  f"cfg2-mf {i2['eval_s'] * 1e3:.2f} ms (factor {i2['factor_MB']:.1f} MB, {i2['flops']:.3e} flop)", flush=True)
        except Exception:
            traceback.print_exc()
    os.environ.pop("HMCMT_MF_LEAF", None)


def stage_xsweep():
    """ordering knobs at cfg2: leaf size x cross-separator size (HMCMT_MF_LEAF, HMCMT_MF_CROSS)"""
    combos = [tuple(int(v) for v in c.split(",")) for c in os.environ.get("XSWEEP", "16,0 16,8 16,13 25,13 36,13 36,0 16,26 36,26 49,13 49,26").split()]
    for leaf, cross in combos:
        os.environ["HMCMT_MF_LEAF"], os.environ["HMCMT_MF_CROSS"] = str(leaf), str(cross)
        try:
            _, _, i2 = plan_eval(200, 100, 30, "mf", nrx=40, reps=5)
            print(f"[xsweep] leaf {leaf} cross {cross}: cfg2-mf {i2['eval_s'] * 1e3:.2f} ms (factor {i2['factor_MB']:.1f} MB, {i2['flops']:.3e} flop, "
                  f"{i2['launches']} launches)", flush=True)
        except Exception:
            traceback.print_exc()
    os.environ.pop("HMCMT_MF_LEAF", None)
    os.environ.pop("HMCMT_MF_CROSS", None)


def stage_fsweep():
    for fs in (48, 64, 80, 96, 112, 128, 144):
        os.environ["HMCMT_MF_FSMALL"] = str(fs)
        try:
            _, _, i4 = plan_eval(800, 300, 8, None, nrx=40, reps=3)
            _, _, i2 = plan_eval(200, 100, 30, "mf", nrx=40, reps=5)
            print(f"[fsweep] fsmall {fs}: cfg4x16 {i4['eval_s'] * 1e3:.2f} ms   cfg2-mf {i2['eval_s'] * 1e3:.2f} ms  launches {i2['launches']}", flush=True)
        except Exception:
            traceback.print_exc()
    os.environ.pop("HMCMT_MF_FSMALL", None)


def stage_crossover():
    for ny, nz, nf in [(48, 40, 8), (96, 56, 11), (76, 52, 12), (120, 60, 10), (150, 80, 20), (200, 100, 30), (400, 100, 10), (104, 104, 10)]:
        try:
            _, _, ib = plan_eval(ny, nz, nf, "band", nrx=10, reps=5)
            _, _, im = plan_eval(ny, nz, nf, "mf", nrx=10, reps=5)
            print(f"[crossover] {ny}x{nz} nf {nf}: band {ib['eval_s'] * 1e3:.3f} ms (T={ib['T']})   mf {im['eval_s'] * 1e3:.3f} ms", flush=True)
        except Exception:
            traceback.print_exc()


if __name__ == "__main__":
    stages = sys.argv[1:] or ["shim", "plan_small", "cfg2", "cfg4"]
    for s in stages:
        print(f"===== {s} =====", flush=True)
        globals()["stage_" + s]()
