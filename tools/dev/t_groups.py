"""Device-resident leapfrog steps of a synthetic workload for several settings of an environment knob (default HMCMT_GROUPS):
ms per step and the final state's difference from the first setting.   python tools/dev/t_groups.py [ny nz nf [KNOB v1 v2 ...]]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hmcmt2d_b200 import api, synthetic  # noqa: E402

ny, nz, nf = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (200, 100, 30)
knob = sys.argv[4] if len(sys.argv) > 4 else "HMCMT_GROUPS"
vals = sys.argv[5:] or ["1", "2", "3", "4"]
mesh, data, inv, prior = synthetic.make_problem(ny, nz, nf)
m0 = synthetic.stress_model(inv)
p0 = np.clip(np.random.default_rng(0).standard_normal(len(m0)), -2.5, 2.5)
ref = None
for v in vals:
    os.environ[knob] = v
    pl = api.Plan(mesh, data, inv, prior)
    pl.set_state(m0, p0, m0)
    pl.leapfrog_steps_device(prior.dt, 3)
    pl.sync()
    best = 1e9
    for _ in range(3):
        pl.timer_start()
        pl.leapfrog_steps_device(prior.dt, 10)
        best = min(best, pl.timer_stop() / 10)
    st = pl.status()
    m, p = pl.get_state()
    if ref is None:
        ref = (m.copy(), p.copy())
    d = max(float(np.abs(m - ref[0]).max()), float(np.abs(p - ref[1]).max()))
    print(f"[{knob}={v}] {ny}x{nz} nf {nf}: {best:.3f} ms/step ({1e3 / best:.1f} steps/s) status {st} |state - first| {d:.2e}", flush=True)
    pl.close()
