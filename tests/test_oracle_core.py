"""Pins the CPU oracle (oracle/ldu_oracle.c) with analytic / independent checks.
The reference ships no golden vectors, so these stand in."""
import numpy as np
import pytest

from conftest import dense_from_ldu


def _case(meshmod, orc, n=6, kind="P", dims=None):
    m = meshmod.hex_mesh(*(dims or (n, n, n)))
    if kind == "P":
        c = meshmod.pressure_laplacian(m)
    else:
        c = meshmod.momentum_matrix(m)
    a = orc.Addr(m.nCells, m.lower, m.upper)
    M = orc.Matrix(a, c["diag"], c["upper"], c["lower"])
    return m, c, a, M


def test_addressing_derived(meshmod, orc):
    m = meshmod.hex_mesh(5, 4, 3)
    a = orc.Addr(m.nCells, m.lower, m.upper)
    os_, ls, lss = a.owner_start(), a.losort(), a.losort_start()
    # faces sorted by owner, upper-triangular
    assert np.all(np.diff(m.lower) >= 0) and np.all(m.lower < m.upper)
    for c in range(m.nCells):
        assert np.all(m.lower[os_[c]:os_[c + 1]] == c)
        fs = ls[lss[c]:lss[c + 1]]
        assert np.all(m.upper[fs] == c)
        assert np.all(np.diff(fs) > 0)  # stable: ascending face index
    assert os_[-1] == m.nFaces and lss[-1] == m.nFaces
    assert m.nFaces == 3 * 5 * 4 * 3 - 4 * 3 - 5 * 3 - 5 * 4


@pytest.mark.parametrize("kind", ["P", "U"])
def test_amul_tmul_vs_dense(meshmod, orc, kind):
    m, c, a, M = _case(meshmod, orc, 5, kind)
    A = dense_from_ldu(m.nCells, m.lower, m.upper, c["diag"], c["upper"], c["lower"])
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, m.nCells)
    np.testing.assert_allclose(M.amul(x), A @ x, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(M.tmul(x), A.T @ x, rtol=1e-13, atol=1e-13)
    y = rng.uniform(-1, 1, m.nCells)
    assert abs(y @ M.amul(x) - M.tmul(y) @ x) < 1e-11  # <y,Ax> = <A^T y,x>
    np.testing.assert_allclose(M.sumA(), A.sum(1), rtol=1e-12, atol=1e-12)
    b = rng.uniform(-1, 1, m.nCells)
    np.testing.assert_allclose(M.residual(x, b), b - A @ x, rtol=1e-12, atol=1e-12)
    Aoff = A - np.diag(np.diag(A))
    np.testing.assert_allclose(M.H(x), -(Aoff @ x), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(M.H1(), -Aoff.sum(1), rtol=1e-12, atol=1e-12)
    lo = c["upper"] if c["lower"] is None else c["lower"]
    np.testing.assert_allclose(M.faceH(x), c["upper"] * x[m.upper] - lo * x[m.lower], rtol=1e-14)


def test_laplacian_eigenpair(meshmod, orc):
    """cos(k pi x) sampled at cell centres is an eigenvector of the Neumann 7-point
    Laplacian (uniform coefficients): A v = -(4/h^2) sin^2(k pi h/2) * h^3 ... per cell."""
    n = 8
    m = meshmod.hex_mesh(n)
    c = meshmod.pressure_laplacian(m, vary=False, pin=False)
    a = orc.Addr(m.nCells, m.lower, m.upper)
    M = orc.Matrix(a, c["diag"], c["upper"], None)
    x = m.cell_centres()[:, 0]
    for k in (1, 2, 3):
        v = np.cos(k * np.pi * x)
        lam = -m.h * 4 * np.sin(k * np.pi * m.h / 2) ** 2  # upper = h, so A = h * (second difference)
        np.testing.assert_allclose(M.amul(v), lam * v, atol=1e-13)


def test_sumdiag_family(meshmod, orc):
    m, c, a, M = _case(meshmod, orc, 4, "U")
    import ctypes as C
    L = orc.lib()
    A = dense_from_ldu(m.nCells, m.lower, m.upper, np.zeros(m.nCells), c["upper"], c["lower"])
    d = np.zeros(m.nCells)
    L.orc_negSumDiag(a.h, orc._d(c["upper"]), orc._d(c["lower"]), orc._d(d))
    # diag[own] -= lower, diag[nei] -= upper  == minus column sums of the off-diagonal part
    np.testing.assert_allclose(d, -A.sum(0), rtol=1e-12, atol=1e-14)
    d2 = np.zeros(m.nCells)
    L.orc_sumDiag(a.h, orc._d(c["upper"]), orc._d(c["lower"]), orc._d(d2))
    np.testing.assert_allclose(d2, A.sum(0), rtol=1e-12, atol=1e-14)
    s = np.zeros(m.nCells)
    L.orc_sumMagOffDiag(a.h, orc._d(c["upper"]), orc._d(c["lower"]), orc._d(s))
    np.testing.assert_allclose(s, np.abs(A).sum(1), rtol=1e-12)


def test_normfactor(meshmod, orc):
    m, c, a, M = _case(meshmod, orc, 5, "P")
    rng = np.random.default_rng(1)
    psi, b = rng.uniform(-1, 1, m.nCells), rng.uniform(-1, 1, m.nCells)
    Apsi = M.amul(psi)
    ref = np.sum(np.abs(Apsi - psi.mean() * M.sumA()) + np.abs(b - psi.mean() * M.sumA())) + 1e-20
    assert abs(M.normFactor(psi, b, Apsi) - ref) / ref < 1e-13


@pytest.mark.parametrize("pre", ["none", "diagonal", "DIC"])
def test_pcg_solves(meshmod, orc, pre):
    m, c, a, M = _case(meshmod, orc, 8, "P")
    A = dense_from_ldu(m.nCells, m.lower, m.upper, c["diag"], c["upper"])
    xs = meshmod.cell_field_global(m, 42)
    b = A @ xs
    psi, perf, hist = M.solve("PCG", pre, np.zeros(m.nCells), b, tolerance=1e-10, maxIter=500)
    assert perf.converged and not perf.singular
    assert perf.solverName.decode() == {"none": "nonePCG", "diagonal": "diagonalPCG", "DIC": "AINVPCG"}[pre]
    np.testing.assert_allclose(psi, xs, atol=1e-6)
    assert len(hist) == perf.nIterations + 1
    assert abs(hist[-1] - perf.finalResidual) == 0 and abs(hist[0] - perf.initialResidual) == 0
    # independent numpy restatement of the same recurrence reproduces the history
    h2 = _numpy_pcg(A, c["diag"], b, pre, len(hist) - 1)
    k = min(25, len(hist))  # CG round-off drift grows with the iteration count
    np.testing.assert_allclose(hist[:k], h2[:k], rtol=1e-9)
    np.testing.assert_allclose(hist, h2, rtol=0.5)


def _numpy_pcg(A, diag, b, pre, nit):
    n = len(b)
    psi = np.zeros(n)
    Aoff = A - np.diag(np.diag(A))
    rD = 1.0 / diag

    def prec(r):
        if pre == "none":
            return r.copy()
        if pre == "diagonal":
            return rD * r
        return rD * (r - Aoff @ (rD * r))
    wA = A @ psi
    rA = b - wA
    sA = A.sum(1)
    nf = np.sum(np.abs(wA - psi.mean() * sA) + np.abs(b - psi.mean() * sA)) + 1e-20
    hist = [np.abs(rA).sum() / nf]
    wArA = 1e20
    pA = None
    for it in range(nit):
        old = wArA
        wA = prec(rA)
        wArA = wA @ rA
        pA = wA.copy() if it == 0 else wA + (wArA / old) * pA
        wA = A @ pA
        alpha = wArA / (wA @ pA)
        psi += alpha * pA
        rA -= alpha * wA
        hist.append(np.abs(rA).sum() / nf)
    return np.array(hist)


def test_pcg_loop_semantics(meshmod, orc):
    """PCG.C:197-205: post-increment => maxIter+1 bodies can run; minIter forces bodies."""
    m, c, a, M = _case(meshmod, orc, 6, "P")
    b = meshmod.cell_field_global(m, 7)
    _, perf, hist = M.solve("PCG", "DIC", np.zeros(m.nCells), b, tolerance=0, maxIter=5)
    assert perf.nIterations == 6 and len(hist) == 7
    _, perf, _ = M.solve("PCG", "DIC", np.zeros(m.nCells), b, tolerance=1e30, maxIter=50, minIter=3)
    assert perf.nIterations == 3
    _, perf, _ = M.solve("PCG", "DIC", np.zeros(m.nCells), b, tolerance=1e30, maxIter=50)
    assert perf.nIterations == 0 and perf.converged
    # relTol
    _, perf, _ = M.solve("PCG", "DIC", np.zeros(m.nCells), b, tolerance=0, relTol=0.1, maxIter=200)
    assert perf.finalResidual < 0.1 * perf.initialResidual and perf.converged


def test_cg_exact_tiny(meshmod, orc):
    """CG terminates in <= n steps on a tiny SPD system (n = 8)."""
    m = meshmod.hex_mesh(2)
    c = meshmod.pressure_laplacian(m)
    a = orc.Addr(m.nCells, m.lower, m.upper)
    M = orc.Matrix(a, c["diag"], c["upper"], None)
    b = np.arange(1.0, 9.0)
    psi, perf, hist = M.solve("PCG", "none", np.zeros(8), b, tolerance=1e-13, maxIter=20)
    assert perf.nIterations <= 9
    np.testing.assert_allclose(M.amul(psi), b, atol=1e-10)


@pytest.mark.parametrize("solver", ["PBiCG", "PBiCGStab"])
@pytest.mark.parametrize("pre", ["none", "diagonal", "DILU"])
def test_asym_solvers(meshmod, orc, solver, pre):
    m, c, a, M = _case(meshmod, orc, 8, "U")
    A = dense_from_ldu(m.nCells, m.lower, m.upper, c["diag"], c["upper"], c["lower"])
    xs = meshmod.cell_field_global(m, 42)
    b = A @ xs
    psi, perf, hist = M.solve(solver, pre, np.zeros(m.nCells), b, tolerance=1e-10, maxIter=300)
    assert perf.converged
    np.testing.assert_allclose(psi, xs, atol=1e-7)


def test_solver_selection_rules(meshmod, orc):
    m, c, a, M = _case(meshmod, orc, 4, "P")
    mu, cu, au, Mu = _case(meshmod, orc, 4, "U")
    b = np.ones(m.nCells)
    with pytest.raises(RuntimeError):
        M.solve("PBiCG", "DILU", b * 0, b)      # asymMatrix table only
    with pytest.raises(RuntimeError):
        Mu.solve("PCG", "DIC", b * 0, b)        # symMatrix table only
    with pytest.raises(RuntimeError):
        M.solve("noSuchSolver", "", b * 0, b)
    with pytest.raises(RuntimeError):
        M.solve("PCG", "FDIC", b * 0, b)
    psi, perf, _ = M.solve("ICCG", "", b * 0, b, tolerance=1e-8)
    assert perf.solverName.decode() == "AINVPCG"
    psi, perf, _ = M.solve("diagonal", "", b * 0, b)
    np.testing.assert_allclose(psi, b / c["diag"])


def test_jacobi_and_smoothsolver(meshmod, orc):
    m, c, a, M = _case(meshmod, orc, 6, "U")
    A = dense_from_ldu(m.nCells, m.lower, m.upper, c["diag"], c["upper"], c["lower"])
    rng = np.random.default_rng(3)
    psi, b = rng.uniform(-1, 1, m.nCells), rng.uniform(-1, 1, m.nCells)
    D = np.diag(A)
    ref = psi.copy()
    for _ in range(3):
        ref = 0.1 * ref + 0.9 / D * (b - (A @ ref - D * ref))
    np.testing.assert_allclose(M.jacobi(psi, b, 3), ref, rtol=1e-12, atol=1e-13)
    xs = meshmod.cell_field_global(m, 5)
    psi2, perf, hist = M.solve("smoothSolver", "GaussSeidel", np.zeros(m.nCells), A @ xs,
                               tolerance=1e-8, maxIter=2000, nSweeps=2)
    assert perf.converged and perf.nIterations % 2 == 0
    np.testing.assert_allclose(psi2, xs, atol=1e-5)
    assert np.all(np.diff(hist) < 0)


def test_pcg_omp_matches_serial(meshmod, orc):
    m, c, a, M = _case(meshmod, orc, 10, "P")
    b = meshmod.cell_field_global(m, 11)
    p1, f1, h1 = M.solve("PCG", "DIC", np.zeros(m.nCells), b, tolerance=1e-9)
    p2, f2 = M.pcg_omp("DIC", np.zeros(m.nCells), b, nThreads=4, tolerance=1e-9)
    assert abs(f1.nIterations - f2.nIterations) <= 1
    np.testing.assert_allclose(p1, p2, atol=1e-7)
    x = meshmod.cell_field_global(m, 12)
    assert np.array_equal(M.amul(x), M.amul_omp(x, 3))


def test_stock_dic_pcg_baseline(meshmod, orc):
    """The stock-OpenFOAM CPU baseline (true DIC) solves the same system (fewer iterations
    than the reference's AINV stand-in)."""
    m, c, a, M = _case(meshmod, orc, 10, "P")
    A = dense_from_ldu(m.nCells, m.lower, m.upper, c["diag"], c["upper"])
    xs = meshmod.cell_field_global(m, 42)
    b = A @ xs
    psi, perf = M.pcg_stock_dic(np.zeros(m.nCells), b, tolerance=1e-10, maxIter=500)
    _, pa, _ = M.solve("PCG", "DIC", np.zeros(m.nCells), b, tolerance=1e-10, maxIter=500)
    assert perf.converged and perf.nIterations <= pa.nIterations
    np.testing.assert_allclose(psi, xs, atol=1e-6)


def _cyclic_case(meshmod, kind="P"):
    """Hex box periodic in x: the x-min and x-max wall faces become a cyclic patch pair."""
    m = meshmod.hex_mesh(6, 5, 4)
    c = meshmod.pressure_laplacian(m, pin=False) if kind == "P" else meshmod.momentum_matrix(m)
    nx, ny, nz = 6, 5, 4
    jk = np.arange(ny * nz)
    lo = (jk * nx).astype(np.int32)                # cells i = 0
    hi = (jk * nx + nx - 1).astype(np.int32)       # cells i = nx-1
    rng = np.random.default_rng(5)
    coeff = rng.uniform(0.5, 1.5, ny * nz) * (1.0 if kind == "P" else 0.05)
    diag = c["diag"].copy()
    if kind == "P":
        np.subtract.at(diag, lo, coeff)
        np.subtract.at(diag, hi, coeff)
        diag[0] *= 2                                # pin (setReference)
        bou = np.concatenate([-coeff, -coeff])
        intc = bou.copy()
    else:
        np.add.at(diag, lo, 1.5 * coeff)
        np.add.at(diag, hi, 0.5 * coeff)
        bou = np.concatenate([0.5 * coeff, 1.5 * coeff])   # distinct both ways: asymmetric coupling
        intc = np.concatenate([1.5 * coeff, 0.5 * coeff])
    ps = np.array([0, ny * nz, 2 * ny * nz], dtype=np.int32)
    fc = np.concatenate([lo, hi]).astype(np.int32)
    nr = np.array([-2, -1], dtype=np.int32)        # patch 0 <-> patch 1
    return m, dict(diag=diag, upper=c["upper"], lower=c["lower"], bou=bou, int=intc), ps, fc, nr, lo, hi


@pytest.mark.parametrize("kind", ["P", "U"])
def test_cyclic_interfaces_vs_dense(meshmod, orc, kind):
    """cyclic coupled patches (cyclicFvPatchField.C:212-231): Amul/Tmul/residual/sumA and the
    solvers against a dense matrix that holds the periodic links explicitly."""
    m, c, ps, fc, nr, lo, hi = _cyclic_case(meshmod, kind)
    a = orc.Addr(m.nCells, m.lower, m.upper, ps, fc, neighbRank=nr)
    M = orc.Matrix(a, c["diag"], c["upper"], c["lower"], c["bou"], c["int"])
    A = dense_from_ldu(m.nCells, m.lower, m.upper, c["diag"], c["upper"], c["lower"])
    n = len(lo)
    A[lo, hi] -= c["bou"][:n]          # Apsi[faceCell] -= bou * psi[partner cell]
    A[hi, lo] -= c["bou"][n:]
    x = np.random.default_rng(1).standard_normal(m.nCells)
    np.testing.assert_allclose(M.amul(x), A @ x, rtol=1e-13, atol=1e-13)
    AT = dense_from_ldu(m.nCells, m.lower, m.upper, c["diag"],
                        c["lower"] if c["lower"] is not None else c["upper"], c["upper"])
    AT[lo, hi] -= c["int"][:n]
    AT[hi, lo] -= c["int"][n:]
    np.testing.assert_allclose(M.tmul(x), AT @ x, rtol=1e-13, atol=1e-13)
    if kind == "P":
        np.testing.assert_allclose(AT, A.T, atol=1e-15)
    b = A @ x
    np.testing.assert_allclose(M.residual(x, b), np.zeros(m.nCells), atol=1e-12)
    np.testing.assert_allclose(M.sumA(), A.sum(axis=1), rtol=1e-12, atol=1e-12)
    for solver, pre in ((("PCG", "DIC"),) if kind == "P" else (("PBiCG", "DILU"), ("PBiCGStab", "DILU"))):
        psi, perf, _ = M.solve(solver, pre, np.zeros(m.nCells), b, tolerance=1e-11, maxIter=500)
        assert perf.converged
        np.testing.assert_allclose(psi, x, atol=1e-7)
    psi, perf, _ = M.solve("smoothSolver", "GaussSeidel", np.zeros(m.nCells), b, tolerance=1e-9, maxIter=5000)
    assert perf.converged
    np.testing.assert_allclose(psi, x, atol=1e-5)
