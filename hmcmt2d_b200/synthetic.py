"""Synthetic problem generator of SURVEY.md section 8(d): mesh pattern copied from the reference's
examples (8 padding cells growing x2 each side of a 200 m core, 7 air layers, 100 m earth layers with the
last 8 growing x2), log-spaced frequencies, receivers on the air/earth interface, 100 Ohm-m background
with a 10 Ohm-m block.  Sizes are totals including padding and air."""
from __future__ import annotations

import numpy as np

from .api import InvDataModel, setupInverseDataModel
from .fileio import HMCPrior, MTData, TensorMesh2D

AIR = np.array([100.0, 300.0, 1000.0, 3000.0, 1e4, 3e4, 1e5])      # listed bottom-up (dprism2d_G96x49.mod:17-18)


def make_mesh(ny: int, nz: int, core_dy: float = 200.0, earth_dz: float = 100.0, npad: int = 8) -> TensorMesh2D:
    if ny <= 2 * npad or nz <= len(AIR) + npad:
        raise ValueError("mesh too small for the padding / air pattern")
    pad = core_dy * 2.0 ** np.arange(1, npad + 1)
    ylen = np.concatenate([pad[::-1], np.full(ny - 2 * npad, core_dy), pad])
    nearth = nz - len(AIR)
    zearth = np.concatenate([np.full(nearth - npad, earth_dz), earth_dz * 2.0 ** np.arange(1, npad + 1)])
    zlen = np.concatenate([AIR[::-1], zearth])
    origin = np.array([pad.sum(), AIR.sum()])
    sigma = np.concatenate([np.full(ny * len(AIR), 1e-8), np.full(ny * nearth, 0.01)])
    return TensorMesh2D(ylen, zlen, AIR.copy(), (ny, nz), origin, sigma)


def true_model(mesh: TensorMesh2D) -> np.ndarray:
    """100 Ohm-m background with a 10 Ohm-m block (dprism-like)."""
    ny, nz = mesh.gridSize
    nair = len(mesh.airLayer)
    sig = np.asarray(mesh.sigma).copy().reshape(nz, ny)
    j0, j1 = int(ny * 0.42), int(ny * 0.58)
    k0, k1 = nair + max(3, (nz - nair) // 10), nair + max(6, (nz - nair) // 4)
    sig[k0:k1, j0:j1] = 0.1
    return sig.reshape(-1)


def make_survey(mesh: TensorMesh2D, nFreq: int, nRx: int = 40, fmax_exp: float = 2.0, fmin_exp: float = -3.0) -> MTData:
    ny = mesh.gridSize[0]
    npad = 8
    core = mesh.yLen[npad:ny - npad].sum()
    freqs = np.logspace(fmax_exp, fmin_exp, nFreq)
    rx = np.stack([np.linspace(0.0, 0.98 * core, nRx), np.zeros(nRx)], axis=1)
    comps = ["ZXY", "ZYX"]
    f, r, c = np.meshgrid(np.arange(1, nFreq + 1), np.arange(1, nRx + 1), np.arange(1, 3), indexing="ij")
    mask = np.ones(nFreq * nRx * 2, dtype=bool)
    return MTData(rx, freqs, "Impedance", comps, r.reshape(-1).astype(np.int64), f.reshape(-1).astype(np.int64),
                  c.reshape(-1).astype(np.int64), mask, True, True)


def halfspace_data(data: MTData, sigma: float = 0.01, noise: float = 0.05, seed: int = 7):
    """Analytic half-space impedances Z = sqrt(i w mu0 / sigma) (ZXY) and -Z (ZYX) with relative noise —
    a cheap observation set for parity tests and the benchmark (data = f(true)(1+0.05 N), err = 0.05|Z|)."""
    rng = np.random.default_rng(seed)
    mu0 = 4e-7 * np.pi
    om = 2 * np.pi * data.freqs[data.freqID - 1]
    z = np.sqrt(1j * om * mu0 / sigma)
    z = np.where(data.dtID == 1, z, -z)
    obs = z * (1.0 + noise * rng.standard_normal(len(z)))
    return obs, noise * np.abs(z)


def make_problem(ny: int, nz: int, nFreq: int, nRx: int = 40, obs=None, err=None, dt: float = 0.03,
                 timestep=(6, 10), rho_bounds=(1.0, 1e4), beta: float = 1.0, **survey_kw):
    """-> (mtMesh, mtData, invParam, hmcprior) exactly as `readstartupFile` would return them."""
    mesh = make_mesh(ny, nz)
    data = make_survey(mesh, nFreq, nRx, **survey_kw)
    if obs is None:
        obs, err = halfspace_data(data)
    inv = setupInverseDataModel(mesh, [1e-8], 1.0 / rho_bounds[1], 1.0 / rho_bounds[0], obs, err)
    prior = HMCPrior(burninsamples=0, totalsamples=10, sigBounds=[1.0 / rho_bounds[1], 1.0 / rho_bounds[0]], dt=dt,
                     timestep=list(timestep), regParam=beta)
    return mesh, data, inv, prior


def stress_model(inv: InvDataModel, seed: int = 1) -> np.ndarray:
    """Evaluation model for timing: ln sigma = ln 0.01 + 0.7 N(0,1) i.i.d. per earth cell."""
    rng = np.random.default_rng(seed)
    return np.log(0.01) + 0.7 * rng.standard_normal(len(inv.strModel))


def true_model_data(data: MTData, pred_true, noise: float = 0.05, seed: int = 7):
    """Observations of SURVEY.md 8(d): data = forward(true model) (1 + 0.05 N), err = 0.05 |Z|.  `pred_true` is the forward
    response of `true_model(mesh)` in the data ordering, computed by whichever engine the caller runs."""
    rng = np.random.default_rng(seed)
    z = np.asarray(pred_true)
    return z * (1.0 + noise * rng.standard_normal(len(z))), noise * np.abs(z)
