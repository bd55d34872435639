"""Row a17 (fvMatrix glue): the oracle's restatement (oracle/fvm_oracle.py) against dense-matrix algebra.
The glue cannot be compiled from the reference (it needs the GeometricField machinery), so these are
analytic checks: every method is compared with what the assembled dense system says it must produce."""
import importlib

import numpy as np
import pytest

import dist_helpers as dh
from oracle import fvm_oracle as fo


def dense(n, lower, upper, diag, up, lo):
    A = np.zeros((n, n))
    A[np.arange(n), np.arange(n)] = diag
    A[lower, upper] = up
    A[upper, lower] = up if lo is None else lo
    return A


def wall_arrays(m):
    bfc = np.concatenate([p.faceCells for p in m.wall_patches()]).astype(np.int32)
    return bfc, np.full(len(bfc), m.h * m.h), np.full(len(bfc), 2.0 / m.h)   # face cells, |Sf|, delta at the wall


def poisson_case(meshmod, orc, dims, seed=3):
    """-laplacian(gamma, p) = s with p fixed on every wall: positive definite, symmetric."""
    m = meshmod.hex_mesh(*dims)
    rng = np.random.default_rng(seed)
    a = orc.Addr(m.nCells, m.lower, m.upper)
    gamma = rng.uniform(0.5, 1.5, m.nFaces)
    lap_upper = m.deltaCoeffs() * (gamma * m.magSf())            # fvm::laplacian: upper, diag = -sum
    upper = -lap_upper                                           # -fvm::laplacian
    diag = np.zeros(m.nCells)
    np.subtract.at(diag, m.lower, upper)
    np.subtract.at(diag, m.upper, upper)
    bfc, magSfb, deltab = wall_arrays(m)
    gb = rng.uniform(0.5, 1.5, len(bfc)) * magSfb
    value = rng.uniform(-1, 1, (len(bfc), 1))
    ic, bc = fo.fixedValue_laplacian_coeffs(gb, deltab, value)
    ic, bc = -ic, -bc                                            # fvMatrix::negate (fvMatrix.C:1738-1750)
    source = rng.uniform(-1, 1, m.nCells) * m.volumes()
    return m, a, dict(diag=diag, upper=upper, lower=None, source=source, bfc=bfc, ic=ic, bc=bc, V=m.volumes())


def momentum_case(meshmod, orc, dims, seed=4):
    """ddt + div(phi) - laplacian(nu) of a vector with fixedValue walls: asymmetric, three components."""
    m = meshmod.hex_mesh(*dims)
    rng = np.random.default_rng(seed)
    a = orc.Addr(m.nCells, m.lower, m.upper)
    c = meshmod.momentum_matrix(m)                               # diag there includes the wall diffusion
    bfc, magSfb, deltab = wall_arrays(m)
    nu = 0.01
    diag = c["diag"].copy()
    np.subtract.at(diag, bfc, nu * m.h * 2.0)                    # take it out again: it enters as internalCoeffs
    value = rng.uniform(-1, 1, (len(bfc), 3))
    ic, bc = fo.fixedValue_laplacian_coeffs(nu * magSfb, deltab, value)
    ic, bc = -ic, -bc
    source = rng.uniform(-1, 1, (m.nCells, 3)) * m.volumes()[:, None]
    return m, a, dict(diag=diag, upper=c["upper"], lower=c["lower"], source=source, bfc=bfc, ic=ic, bc=bc,
                      V=m.volumes())


def full_system(m, d, k=0):
    """dense matrix and right-hand side of component k with the boundary folded in"""
    A = dense(m.nCells, m.lower, m.upper, d["diag"], d["upper"], d["lower"])
    rhs = np.array(d["source"], float).reshape(m.nCells, -1)[:, k].copy()
    np.add.at(A, (d["bfc"], d["bfc"]), d["ic"][:, k])
    np.add.at(rhs, d["bfc"], d["bc"][:, k])
    return A, rhs


def make(orc, a, d, nc, psi=None, **kw):
    n = a.nCells
    return fo.FvMatrix(orc, a, nc, d["diag"], d["upper"], d["lower"], d["source"],
                       np.zeros((n, nc)) if psi is None else psi, d["V"], d["bfc"], d["ic"], d["bc"], **kw)


def test_scalar_solve_A_H_flux_residual(meshmod, orc):
    m, a, d = poisson_case(meshmod, orc, (7, 6, 5))
    A, rhs = full_system(m, d)
    exact = np.linalg.solve(A, rhs)
    fm = make(orc, a, d, 1)
    psi, perfs, _ = fm.solve("PCG", "DIC", tolerance=1e-13, maxIter=500)
    assert perfs[0].converged
    np.testing.assert_allclose(psi[:, 0], exact, rtol=0, atol=1e-10)
    assert np.array_equal(fm.diag, d["diag"])                    # the saved diagonal is restored (:187)
    # at the solution A*psi = H (both divided by V) and the residual vanishes
    fs = make(orc, a, d, 1, exact[:, None])
    np.testing.assert_allclose(fs.A() * exact, fs.H()[:, 0], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(fs.residual(), 0, atol=1e-10)
    # away from it: residual = rhs - A psi, and V*(A psi - H) = -residual
    x = np.random.default_rng(0).uniform(-1, 1, m.nCells)
    fx = make(orc, a, d, 1, x[:, None])
    np.testing.assert_allclose(fx.residual(), rhs - A @ x, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(d["V"] * (fx.A() * x - fx.H()[:, 0]), -(rhs - A @ x), rtol=1e-10, atol=1e-11)
    # flux: summed over the faces of a cell (owner +, neighbour -, boundary +) it is the matrix applied to psi
    internal, boundary, coupled = fx.flux()
    assert coupled.shape == (0, 1)
    div = np.asarray(orc.surface_integrate(a, internal[:, 0], d["bfc"], boundary[:, 0], d["V"], 1, False, -1))
    np.testing.assert_allclose(div, A @ x - (rhs - d["source"]), rtol=1e-10, atol=1e-11)


def test_vector_solve_segregated_and_H(meshmod, orc):
    m, a, d = momentum_case(meshmod, orc, (6, 5, 4))
    fm = make(orc, a, d, 3)
    psi, perfs, _ = fm.solve("PBiCG", "DILU", tolerance=1e-13, maxIter=500)
    assert len(perfs) == 3 and all(p.converged for p in perfs)
    x = np.random.default_rng(1).uniform(-1, 1, (m.nCells, 3))
    fx = make(orc, a, d, 3, x)
    H, Aphi = fx.H(), fx.A()
    for k in range(3):
        A, rhs = full_system(m, d, k)
        np.testing.assert_allclose(psi[:, k], np.linalg.solve(A, rhs), rtol=0, atol=1e-10)
        # internalCoeffs are equal in all components here, so D() is the folded diagonal of every component
        np.testing.assert_allclose(d["V"] * (Aphi * x[:, k] - H[:, k]), A @ x[:, k] - rhs, rtol=1e-10, atol=1e-11)
    # ... and the stock boundary-diagonal term is rounding noise ((x + x + x)/3 - x)
    np.testing.assert_allclose(fx.H(boundaryDiagInH=True), H, rtol=1e-12, atol=1e-12)
    # component-dependent internal coefficients: the reference's H drops the term stock OpenFOAM keeps
    d2 = dict(d)
    d2["ic"] = d["ic"] * np.array([1.0, 2.0, 3.0])
    f2 = make(orc, a, d2, 3, x)
    assert not np.allclose(f2.H(boundaryDiagInH=True), f2.H())
    Dav = f2.D()
    for k in range(3):
        A, rhs = full_system(m, d2, k)
        stock = f2.H(boundaryDiagInH=True)[:, k]
        np.testing.assert_allclose(Dav * x[:, k] - d["V"] * stock, A @ x[:, k] - rhs, rtol=1e-10, atol=1e-11)


def test_set_reference_and_relax(meshmod, orc):
    m, a, d = momentum_case(meshmod, orc, (5, 5, 4))
    A, rhs = full_system(m, d, 0)
    exact = np.stack([np.linalg.solve(*full_system(m, d, k)) for k in range(3)], axis=1)
    fm = make(orc, a, d, 3, exact)
    D0, S0 = fm.diag.copy(), fm.source.copy()
    fm.relax(1.0)                                                # dominant matrix, alpha 1: nothing changes
    np.testing.assert_allclose(fm.diag, D0, rtol=1e-14)
    np.testing.assert_allclose(fm.source, S0, rtol=1e-12, atol=1e-15)
    fm = make(orc, a, d, 3, exact)
    fm.relax(0.7)
    assert np.all(fm.diag > D0)
    # the relaxed system has the same solution when psi already is the solution
    d2 = dict(d, diag=fm.diag, source=fm.source)
    for k in range(3):
        A2, rhs2 = full_system(m, d2, k)
        np.testing.assert_allclose(np.linalg.solve(A2, rhs2), exact[:, k], rtol=0, atol=1e-9)
    # relax restores diagonal dominance: an off-diagonal row sum larger than the diagonal lifts the diagonal
    d3 = dict(d, diag=d["diag"] * 1e-3)
    f3 = make(orc, a, d3, 3, exact)
    f3.relax(1.0)
    sumOff = np.zeros(m.nCells)
    np.add.at(sumOff, m.lower, np.abs(d["upper"]))
    np.add.at(sumOff, m.upper, np.abs(d["lower"]))
    folded = f3.diag.copy()
    np.add.at(folded, d["bfc"], d["ic"][:, 0])
    assert np.all(folded >= sumOff * (1 - 1e-12))
    # setReference (fvMatrix.C:965-983)
    f4 = make(orc, a, d, 3)
    f4.setReference(7, [1.0, -2.0, 0.5])
    assert f4.diag[7] == 2 * d["diag"][7]
    np.testing.assert_array_equal(f4.source[7], d["source"][7] + d["diag"][7] * np.array([1.0, -2.0, 0.5]))
    f4.setReference(-1, [1.0, 1.0, 1.0])                         # no reference cell on this rank
    assert f4.diag[7] == 2 * d["diag"][7]


def decomposed_global(meshmod, orc, n):
    """-laplacian(gamma, U) of a vector with fixedValue walls on the single n^3 domain (the data every rank cuts from)"""
    gm = meshmod.hex_mesh(n)
    rng = np.random.default_rng(9)
    gamma = rng.uniform(0.5, 1.5, gm.nFaces)
    upper = -(gm.deltaCoeffs() * (gamma * gm.magSf()))
    diag = np.zeros(gm.nCells)
    np.subtract.at(diag, gm.lower, upper)
    np.subtract.at(diag, gm.upper, upper)
    bfc, magSfb, deltab = wall_arrays(gm)
    value = rng.uniform(-1, 1, (len(bfc), 3))
    ic, bc = fo.fixedValue_laplacian_coeffs(magSfb, deltab, value)
    gd = dict(diag=diag, upper=upper, lower=None, source=rng.uniform(-1, 1, (gm.nCells, 3)) * gm.h ** 3, bfc=bfc,
              ic=-ic, bc=-bc, V=gm.volumes())
    x = rng.uniform(-1, 1, (gm.nCells, 3))
    fkey = {(int(l), int(u)): f for f, (l, u) in enumerate(zip(gm.lower, gm.upper))}
    wkey, off = {}, 0
    for p in gm.wall_patches():
        for i, c in enumerate(p.faceCells):
            wkey[(int(c), p.name)] = off + i
        off += len(p.faceCells)
    return dict(gm=gm, gd=gd, x=x, upper=upper, fkey=fkey, wkey=wkey, n=n)


def decomposed_rank(meshmod, G, nR, r):
    """rank r's cut of decomposed_global: (mesh, fvMatrix arrays, coupled coefficient per coupled face).  A coupled face
    carries internalCoeffs = -upper_f (what the missing neighbour row would have put on the diagonal) and
    boundaryCoeffs = -upper_f (Amul subtracts boundaryCoeffs*psi_nbr)."""
    gm, gd, upper, fkey, wkey = G["gm"], G["gd"], G["upper"], G["fkey"], G["wkey"]
    m = meshmod.decompose(G["n"], nR, r)
    cg = m.cellGlobal
    gl = np.array([fkey[(int(cg[l]), int(cg[u]))] for l, u in zip(m.lower, m.upper)])
    cou = []
    for p in m.coupled_patches():
        mine, theirs = cg[p.faceCells], p.nbrGlobalCells
        cou.append(np.array([upper[fkey[(min(int(i), int(j)), max(int(i), int(j)))]] for i, j in zip(mine, theirs)]))
    cou = np.concatenate(cou)
    d0 = np.zeros(m.nCells)
    np.subtract.at(d0, m.lower, upper[gl])
    np.subtract.at(d0, m.upper, upper[gl])
    wsel = np.array([wkey[(int(cg[c]), p.name)] for p in m.wall_patches() for c in p.faceCells], int)
    wb = np.concatenate([p.faceCells for p in m.wall_patches()]).astype(np.int32)
    d = dict(diag=d0, upper=upper[gl], lower=None, source=gd["source"][cg], bfc=wb, ic=gd["ic"][wsel],
             bc=gd["bc"][wsel], V=gm.volumes()[cg])
    return m, d, cou


def oracle_rank_results(meshmod, orc, G, nR, ctl):
    """what every rank's oracle FvMatrix produces (threads): list of dicts, one per rank"""
    ex = dh.ThreadExchange(nR)
    x, n = G["x"], G["n"]

    def rank_fn(r):
        m, d, cou = decomposed_rank(meshmod, G, nR, r)
        cg = m.cellGlobal
        ps, fc = m.patch_start_facecells()
        a = orc.Addr(m.nCells, m.lower, m.upper, ps, fc, neighbRank=[p.neighbRank for p in m.coupled_patches()])
        kw = dict(couInt=-cou, couBou=-cou, comm=ex.comm(orc, r, m, n ** 3))
        f = make(orc, a, d, 3, x[cg], **kw)
        pnf = f.patchNeighbourField()
        H, Aphi = f.H(), f.A()
        internal, boundary, coupled = f.flux()
        fz = make(orc, a, d, 3, **kw)
        psi, perfs, _ = fz.solve("PCG", "DIC", **ctl)
        fsc = make(orc, a, dict(d, source=d["source"][:, 0], ic=d["ic"][:, :1], bc=d["bc"][:, :1]), 1, x[cg][:, :1], **kw)
        return dict(cg=cg, H=H, A=Aphi, psi=psi, res=fsc.residual(), internal=internal, boundary=boundary,
                    coupled=coupled, pnf=pnf, conv=[p.converged for p in perfs], nIter=[p.nIterations for p in perfs],
                    cou=cou, m=m)
    return dh.run_threads(nR, rank_fn)


@pytest.mark.parametrize("nR", [2, 4])
def test_decomposed_glue_matches_single_domain(meshmod, orc, nR):
    """Coupled patches: solve, H, flux and residual of the decomposed case reproduce the single domain."""
    G = decomposed_global(meshmod, orc, 8)
    gm, gd, x = G["gm"], G["gd"], G["x"]
    ga = orc.Addr(gm.nCells, gm.lower, gm.upper)
    gf = make(orc, ga, gd, 3, x)
    gH, gA = gf.H(), gf.A()
    gpsi, _, _ = make(orc, ga, gd, 3).solve("PCG", "DIC", tolerance=1e-13, maxIter=800)
    gs = make(orc, ga, dict(gd, source=gd["source"][:, 0], ic=gd["ic"][:, :1], bc=gd["bc"][:, :1]), 1, x[:, :1])
    gres = gs.residual()
    for R in oracle_rank_results(meshmod, orc, G, nR, dict(tolerance=1e-13, maxIter=800)):
        cg, m, cou = R["cg"], R["m"], R["cou"]
        assert all(R["conv"])
        np.testing.assert_allclose(R["H"], gH[cg], rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(R["A"], gA[cg], rtol=1e-13)
        np.testing.assert_allclose(R["psi"], gpsi[cg], rtol=0, atol=1e-9)
        # fvScalarMatrix.C:195-240 counts the coupled neighbour term twice: once inside lduMatrix::residual (the
        # interface update) and once more in addBoundarySource(res) whose `couples` defaults to true -- as written
        extra = np.zeros(len(cg))
        nbr = np.concatenate([p.nbrGlobalCells for p in m.coupled_patches()])
        np.add.at(extra, np.concatenate([p.faceCells for p in m.coupled_patches()]), (-cou) * x[nbr, 0])
        np.testing.assert_allclose(R["res"], gres[cg] + extra, rtol=1e-11, atol=1e-11)
        assert np.array_equal(R["pnf"], x[nbr])
        # flux through a processor face: internalCoeffs*psi_i - boundaryCoeffs*psi_j
        k = 0
        for p in m.coupled_patches():
            for i, j in zip(m.cellGlobal[p.faceCells], p.nbrGlobalCells):
                np.testing.assert_allclose(R["coupled"][k], (-cou[k]) * x[i] - (-cou[k]) * x[j], rtol=1e-13)
                k += 1


def test_oracle_reproduces_fvm_golden(meshmod, orc):
    """tests/golden/fvm_golden.npz (self-generated, make_fvm_golden.py): the glue and two icoFoam steps, bit for bit"""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_fvm_golden as mg
    golden = np.load(os.path.join(here, "golden", "fvm_golden.npz"))
    fresh = mg.generate(meshmod, orc)
    assert sorted(fresh) == sorted(golden.files)
    for k in golden.files:
        assert np.array_equal(fresh[k], golden[k]), k
