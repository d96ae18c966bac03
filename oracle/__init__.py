"""CPU oracle for the HMCMT2D forward + adjoint-gradient hot path.

TEST INFRASTRUCTURE ONLY.  This package is a literal numpy/scipy restatement of the
reference's Julia algorithm (CUG-EMI/HMCMT2D, `HMCMT/src/**`), each function citing the
reference file:line it follows.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product path
(`hmcmt2d_b200`) never does and fails loudly when its CUDA library is missing.

PARITY STATUS: **parity unpinned by the reference** — the reference ships no golden
vectors or tests for this path (only `MUMPS/test/*` solver-residual tests) and cannot
be executed in the build container (no Julia, MUMPS binary absent).  The oracle is
therefore pinned by known-answer tests instead (tests/test_oracle_*.py): analytic
half-space / layered-earth impedances, J^T v == (explicit J)^T v where the explicit J is
restated from a *different* reference file (`compJacMat.jl`), finite-difference checks
on cells where the reference's own boundary approximations do not bite, and the
MUMPS-test residual criteria at the solver boundary.

Indices are 0-based here; every place where the reference's 1-based index enters a
result (DOF numbering, CSC pattern, file formats) is converted explicitly and noted.
The sparse direct solves use SciPy SuperLU in place of UMFPACK/MUMPS (third-party,
un-vendored in the reference: SuiteSparse 7.2.1 `HMCMT/Manifest.toml:436-439`, MUMPS
binary `.MISSING_LARGE_BLOBS`); cross-ordering agreement is ~5e-12, inside the 1e-9
budget.
"""
