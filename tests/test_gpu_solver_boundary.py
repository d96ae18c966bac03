"""Level-1 ABI (MUMPS shim) on the GPU — mirrors the reference's own solver-boundary tests
MUMPS/test/testDivGrad.jl (:19,32-33,45-46,58-59) and testTwoSystem.jl (:35,43): relative residual
< 1e-14, result eltype, several live factorisations — at the reference's own sizes (32x32x16; 24x23x25 and 34x32x36), which
the library orders by nested dissection itself, plus a small grid that fits the register-window band kernel
(b = n1*n2 <= 104), and the matrices the unmodified reference would actually pass: y-fastest Aii of the MT systems."""
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


@pytest.mark.parametrize("dims", [(10, 10, 16), (32, 32, 16)])
def test_divgrad_real_and_complex(dims):
    """testDivGrad.jl:9-59 (second size = the reference's)."""
    from hmcmt2d_b200 import lib
    rng = np.random.default_rng(0)
    A = getDivGrad(*dims)
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


@pytest.mark.parametrize("d1,d2", [((8, 9, 11), (10, 9, 12)), ((24, 23, 25), (34, 32, 36))])
def test_two_live_factorisations(d1, d2):
    """testTwoSystem.jl:23-47 (second pair = the reference's sizes)."""
    from hmcmt2d_b200 import lib
    rng = np.random.default_rng(1)
    A = getDivGrad(*d1)
    A2 = getDivGrad(*d2)
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
    rhs = np.arange(500.0)
    assert relres(wide, lib.solveMUMPS(wide, rhs, 1), rhs) < 1e-14      # any bandwidth: the library reorders
    with pytest.raises(ValueError):
        lib.factorMUMPS(sp.csc_matrix(np.ones((3, 4))), 1)


def test_reference_ordered_mt_systems():
    """What the unmodified reference hands to factorMUMPS: Aii in ITS numbering (y fastest, MT2DFwdSolver.jl:232-234), i.e.
    half-bandwidth ny-1 = 199 at cfg2 — exported by hmcmt_export_system, solved through the Level-1 symbols, compared with the
    fused path's own field.  Two frequencies x two modes keep four factorisations of the same pattern alive (symbolic cache)."""
    from hmcmt2d_b200 import api, lib, synthetic
    mesh, data, inv, prior = synthetic.make_problem(200, 100, 2, nRx=10)
    m = synthetic.stress_model(inv)
    pl = api.Plan(mesh, data, inv, prior)
    pred, ex, hx = pl.forward(m=m)
    N = pl.info(0)
    ny, nz = mesh.gridSize
    handles = []
    for mode, fld in ((0, ex), (1, hx)):
        for f in range(2):
            colptr, rowval, nzval, rhs, bc = pl.export_system(mode, f)
            A = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(N, N))
            assert np.abs(A.indices - np.repeat(np.arange(N), np.diff(A.indptr))).max() == ny - 1      # b = 199 > 104
            F = lib.factorMUMPS(A, 1)
            handles.append(F)
            x = lib.applyMUMPS(F, rhs)
            # these matrices are badly scaled (cond_1 up to 1e13 in TM): the yardstick is what a pivoting CPU solver achieves
            import scipy.sparse.linalg as spla
            ref = relres(A, spla.splu(A).solve(rhs), rhs)
            assert relres(A, x, rhs) < max(1e-13, 10 * ref), (mode, f, ref)
            interior = fld[0, f].reshape(nz + 1, ny + 1)[1:-1, 1:-1].reshape(-1)                          # y fastest
            assert np.abs(x - interior).max() / np.abs(interior).max() < 1e-10
    for F in handles:
        lib.destroyMUMPS(F)
    pl.close()
