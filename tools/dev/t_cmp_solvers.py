import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
from hmcmt2d_b200 import api, synthetic
for ny, nz, nf in [(40, 30, 6), (20, 16, 4)]:
    mesh, data, inv, prior = synthetic.make_problem(ny, nz, nf)
    m0 = synthetic.stress_model(inv)
    p0 = np.clip(np.random.default_rng(0).standard_normal(len(m0)), -2.5, 2.5)
    res = {}
    for solver in ("band", "mf"):
        os.environ["HMCMT_SOLVER"] = solver
        pl = api.Plan(mesh, data, inv, prior)
        pred, phi, g = pl.forward_gradient(m0)
        traj = []
        pl.set_state(m0, p0, m0)
        for k in range(12):
            pl.leapfrog_steps_device(prior.dt, 1)
            m, p = pl.get_state()
            traj.append((m.copy(), p.copy()))
        res[solver] = (pred, phi, g, traj)
        pl.close()
    a, b = res["band"], res["mf"]
    print(ny, nz, "single evaluation: pred", float((np.abs(a[0] - b[0]) / np.abs(a[0])).max()), "phi", abs(a[1][0] - b[1][0]) / abs(a[1][0]),
          "grad", float(np.abs(a[2] - b[2]).max() / np.abs(a[2]).max()), "|g|max", float(np.abs(a[2]).max()))
    print("   per-step state difference:", ["%.1e" % max(np.abs(x[0] - y[0]).max(), np.abs(x[1] - y[1]).max()) for x, y in zip(a[3], b[3])])
    print("   |p|max per step (band):", ["%.1e" % np.abs(x[1]).max() for x in a[3]])
