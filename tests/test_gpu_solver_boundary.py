"""Level-1 ABI (MUMPS shim) on the GPU — mirrors the reference's own solver-boundary tests
MUMPS/test/testDivGrad.jl (:19,32-33,45-46,58-59) and testTwoSystem.jl (:35,43): relative residual
< 1e-14, result eltype, several live factorisations.  The div-grad grid is sized so that its
half-bandwidth fits the register-window kernel (b = n1*n2 <= 104); larger bandwidths: test_gpu_bigband.py."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def ddx(n):
    return sp.diags([-np.ones(n), np.ones(n)], [0, 1], shape=(n, n + 1))


def getDivGrad(n1, n2, n3):
    """MUMPS/test/getDivGrad.jl:3-13."""
    I = sp.identity
    D1 = sp.kron(I(n3), sp.kron(I(n2), ddx(n1)))
    D2 = sp.kron(I(n3), sp.kron(ddx(n2), I(n1)))
    D3 = sp.kron(ddx(n3), sp.kron(I(n2), I(n1)))
    Div = sp.hstack([D1, D2, D3])
    return (Div @ Div.T).tocsc()


def relres(A, x, b):
    if b.ndim == 1:
        return np.linalg.norm(A @ x - b) / np.linalg.norm(b)
    return max(np.linalg.norm(A @ x[:, i] - b[:, i]) / np.linalg.norm(b[:, i]) for i in range(b.shape[1]))


def test_divgrad_real_and_complex():
    from hmcmt2d_b200 import lib
    rng = np.random.default_rng(0)
    A = getDivGrad(10, 10, 16)
    n = A.shape[0]
    rhs = rng.standard_normal(n)
    x = lib.solveMUMPS(A, rhs, 1)
    assert x.dtype == np.float64 and relres(A, x, rhs) < 1e-14
    rhs = rng.standard_normal((n, 10))
    x = lib.solveMUMPS(A, rhs, 1)
    assert x.dtype == np.float64 and relres(A, x, rhs) < 1e-14
    Ac = (A + 1j * sp.diags(rng.random(n))).tocsc()
    rhs = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x = lib.solveMUMPS(Ac, rhs, 1)
    assert x.dtype == np.complex128 and relres(Ac, x, rhs) < 1e-14
    rhs = rng.standard_normal((n, 10)) + 1j * rng.standard_normal((n, 10))
    x = lib.solveMUMPS(Ac, rhs, 2)
    assert x.dtype == np.complex128 and relres(Ac, x, rhs) < 1e-14


def test_two_live_factorisations():
    from hmcmt2d_b200 import lib
    rng = np.random.default_rng(1)
    A = getDivGrad(8, 9, 11)
    A2 = getDivGrad(10, 9, 12)
    A = (A + 1j * sp.diags(rng.random(A.shape[0]))).tocsc()
    rhs = rng.standard_normal((A.shape[0], 10)) + 1j * rng.standard_normal((A.shape[0], 10))
    rhs2 = rng.standard_normal((A2.shape[0], 10))
    F1 = lib.factorMUMPS(A, 1)
    F2 = lib.factorMUMPS(A2, 1)
    x = lib.applyMUMPS(F1, rhs)
    x2 = lib.applyMUMPS(F2, rhs2)
    assert relres(A, x, rhs) < 1e-14 and relres(A2, x2, rhs2) < 1e-14
    assert x2.dtype == np.float64
    lib.destroyMUMPS(F1)
    lib.destroyMUMPS(F2)
    assert F1.ptr == -1 and F2.n == -1
    with pytest.raises(lib.HmcmtError):
        lib.applyMUMPS(lib.MUMPSfactorization(12345, A.shape[0], True), rhs)


def test_band_edge_sizes_and_padding():
    """Every tile-window size T in {2..14} and N not a multiple of 8 (identity padding past the end)."""
    from hmcmt2d_b200 import lib
    rng = np.random.default_rng(2)
    for nl, nf in [(5, 3), (9, 9), (7, 24), (6, 37), (5, 50), (4, 70), (4, 88), (3, 99), (3, 104)]:
        N = nl * nf
        d = 4 + rng.random(N) + 1j * rng.random(N)
        e1, e2 = -rng.random(N), -rng.random(N)
        e1[np.arange(N) % nf == 0] = 0
        A = sp.diags([d, e1[1:], e1[1:], e2[nf:], e2[nf:]], [0, -1, 1, -nf, nf], format="csc")
        rhs = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        assert relres(A, lib.solveMUMPS(A, rhs, 1), rhs) < 1e-14, (nl, nf)


def test_error_codes():
    from hmcmt2d_b200 import lib
    n = 40
    A = sp.diags([np.full(n, 4.0), np.full(n - 1, -1.0), np.full(n - 1, -1.0)], [0, 1, -1], format="csc")
    with pytest.raises(lib.HmcmtError) as e:
        lib.factorMUMPS(A, 0)                      # unsymmetric factorisation is not provided
    assert e.value.code == -3
    zero = sp.csc_matrix((np.zeros(n), (np.arange(n), np.arange(n))), shape=(n, n))      # explicit zeros on the diagonal
    with pytest.raises(lib.HmcmtError) as e:
        lib.factorMUMPS(zero, 1)
    assert e.value.code == -10                     # "Numerically singular matrix" (MUMPSfuncs.jl:62-63)
    wide = sp.diags([np.full(500, 4.0), np.full(100, -1.0), np.full(100, -1.0)], [0, 400, -400], format="csc")
    with pytest.raises(lib.HmcmtError) as e:
        lib.factorMUMPS(wide, 1)                   # half-bandwidth 400 > 320: refused, never a CPU fallback
    assert e.value.code == -3
    with pytest.raises(ValueError):
        lib.factorMUMPS(sp.csc_matrix(np.ones((3, 4))), 1)
