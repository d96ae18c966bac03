"""Regenerates tests/golden/tiny_oracle.npz: the CPU oracle's outputs (predicted data, misfit, gradient, a 2-step leapfrog
trajectory) on the seeded 8x8-cell problem of tests/helpers.tiny_problem.  The reference itself cannot run here (Julia and
the MUMPS binary are absent), so this fixture pins the ORACLE against accidental change and gives the GPU tests a committed
vector to compare with; it is not a reference output.

    python tests/golden/make_oracle_fixture.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sampler as osamp                     # noqa: E402
from tests.helpers import tiny_problem                  # noqa: E402


def compute():
    mesh, data, inv, prior = tiny_problem(seed=3)
    m = inv.strModel.copy()
    pred, phi, g = osamp.compDataGradient(mesh, data, inv, prior)
    p0 = np.clip(np.random.default_rng(0).standard_normal(len(m)), -2.5, 2.5)
    m2, p2 = osamp.proposeLeapfrog(m.copy(), p0.copy(), mesh, data, inv, prior, 2)
    return dict(m=m, pred=pred, phi=np.float64(phi), g=g, p0=p0, m2=m2, p2=p2)


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_oracle.npz")
    np.savez(out, **compute())
    print("wrote", out)
