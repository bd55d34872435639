"""Pins the oracle's GAMG (pair agglomeration, coarse addressing, V-cycle) and FV face-sum
restatements with independent checks (the reference has no fixtures for them)."""
import numpy as np
import pytest

from conftest import dense_from_ldu


def _gamg_case(meshmod, orc, dims, kind="P"):
    m = meshmod.hex_mesh(*dims)
    c = meshmod.pressure_laplacian(m) if kind == "P" else meshmod.momentum_matrix(m)
    a = orc.Addr(m.nCells, m.lower, m.upper)
    M = orc.Matrix(a, c["diag"], c["upper"], c["lower"])
    g = orc.Gamg(a, meshmod.face_area_pair_weights(m), nCellsInCoarsestLevel=10)
    return m, c, a, M, g


def test_pair_agglomeration_structure(meshmod, orc):
    m, c, a, M, g = _gamg_case(meshmod, orc, (12, 10, 8))
    assert g.nLevels >= 3
    n = m.nCells
    fine_l, fine_u = m.lower, m.upper
    for lev in range(g.nLevels):
        r = g.restrict_addr(lev)
        nc = g.ncells(lev)
        assert r.min() == 0 and r.max() == nc - 1 and len(r) == n
        cnt = np.bincount(r, minlength=nc)
        assert cnt.min() >= 1 and cnt.max() <= 4       # pairs (+ attached leftovers)
        assert nc >= 10 and nc <= 0.75 * n + 1
        la = g.level_addr(lev)
        cl, cu = la.lower(), la.upper()
        assert np.all(cl < cu) and np.all(np.diff(cl) >= 0)   # upper-triangular, owner-grouped
        fr, fl = g.face_restrict_addr(lev), g.face_flip(lev)
        same = r[fine_l] == r[fine_u]
        assert np.array_equal(fr < 0, same)
        assert np.array_equal(-1 - fr[same], r[fine_l][same])
        k = ~same
        own, nei = np.minimum(r[fine_l][k], r[fine_u][k]), np.maximum(r[fine_l][k], r[fine_u][k])
        assert np.array_equal(cl[fr[k]], own) and np.array_equal(cu[fr[k]], nei)
        assert np.array_equal(fl[k].astype(bool), r[fine_u][k] == own)
        assert len(set(zip(cl.tolist(), cu.tolist()))) == len(cl)  # unique coarse faces
        fine_l, fine_u, n = cl, cu, nc
    # weights 1 / 1.01 / 1.02 pair z-neighbours first on the finest level (faceAreaPair tie-break)
    r0 = g.restrict_addr(0)
    zf = m.faceDir == 2
    assert (r0[m.lower[zf]] == r0[m.upper[zf]]).mean() > 0.45
    # forward_ toggles once per agglomerate() call, including the rejected last one
    assert g.forward == (1 if (g.nLevels + 1) % 2 == 0 else 0)


def test_forward_flag_persists(meshmod, orc):
    m = meshmod.hex_mesh(7, 5, 3)  # odd sizes: forward and backward sweeps pair differently
    a = orc.Addr(m.nCells, m.lower, m.upper)
    w = meshmod.face_area_pair_weights(m)
    g1 = orc.Gamg(a, w, 10, forward=1)
    g2 = orc.Gamg(a, w, 10, forward=0)
    assert not np.array_equal(g1.restrict_addr(0), g2.restrict_addr(0))
    assert g1.nLevels == g2.nLevels


@pytest.mark.parametrize("kind,kw", [("P", {}), ("P", dict(nPreSweeps=1)), ("P", dict(interpolateCorrection=1)),
                                     ("U", {}), ("U", dict(scaleCorrection=1, nPreSweeps=2, nFinestSweeps=1))])
def test_gamg_converges(meshmod, orc, kind, kw):
    m, c, a, M, g = _gamg_case(meshmod, orc, (12, 12, 12), kind)
    A = dense_from_ldu(m.nCells, m.lower, m.upper, c["diag"], c["upper"], c["lower"])
    xs = meshmod.cell_field_global(m, 42)
    b = A @ xs
    psi, perf, hist = g.solve(M, "GaussSeidel", np.zeros(m.nCells), b, tolerance=1e-8, maxIter=100, **kw)
    assert perf.converged and perf.nIterations < 60
    assert np.all(np.diff(hist) < 0)
    np.testing.assert_allclose(psi, xs, atol=1e-5)
    assert len(hist) == perf.nIterations + 1


def test_merge_levels_compose_pair_steps(meshmod, orc):
    """mergeLevels 2 (combineLevels, GAMGAgglomerateLduAddressing.C:606-765): level k is the
    composition of pair steps 2k and 2k+1 of the mergeLevels-1 hierarchy (the pairing itself
    sees the same addressing and weights in the same order)."""
    m = meshmod.hex_mesh(12, 10, 8)
    a = orc.Addr(m.nCells, m.lower, m.upper)
    w = meshmod.face_area_pair_weights(m)
    g1 = orc.Gamg(a, w, 10, mergeLevels=1)
    g2 = orc.Gamg(a, w, 10, mergeLevels=2)
    assert g2.nLevels == (g1.nLevels + 1) // 2
    fine_l, fine_u = m.lower, m.upper
    for k in range(g2.nLevels):
        r = g1.restrict_addr(2 * k)
        if 2 * k + 1 < g1.nLevels:
            r = g1.restrict_addr(2 * k + 1)[r]
            last = 2 * k + 1
        else:
            last = 2 * k
        assert np.array_equal(g2.restrict_addr(k), r)
        assert g2.ncells(k) == g1.ncells(last) and g2.nfaces(k) == g1.nfaces(last)
        la2, la1 = g2.level_addr(k), g1.level_addr(last)
        assert np.array_equal(la2.lower(), la1.lower()) and np.array_equal(la2.upper(), la1.upper())
        # composed face map is consistent with the composed cell map
        fr = g2.face_restrict_addr(k)
        same = r[fine_l] == r[fine_u]
        assert np.array_equal(fr < 0, same)
        assert np.array_equal(-1 - fr[same], r[fine_l][same])
        cl, cu = la2.lower(), la2.upper()
        kk = ~same
        assert np.array_equal(cl[fr[kk]], np.minimum(r[fine_l][kk], r[fine_u][kk]))
        assert np.array_equal(cu[fr[kk]], np.maximum(r[fine_l][kk], r[fine_u][kk]))
        fine_l, fine_u = cl, cu
    assert g2.forward == g1.forward


@pytest.mark.parametrize("kind", ["P", "U"])
def test_gamg_merge_levels_and_krylov_coarsest_converge(meshmod, orc, kind):
    m = meshmod.hex_mesh(12, 12, 12)
    c = meshmod.pressure_laplacian(m) if kind == "P" else meshmod.momentum_matrix(m)
    a = orc.Addr(m.nCells, m.lower, m.upper)
    M = orc.Matrix(a, c["diag"], c["upper"], c["lower"])
    A = dense_from_ldu(m.nCells, m.lower, m.upper, c["diag"], c["upper"], c["lower"])
    xs = meshmod.cell_field_global(m, 42)
    b = A @ xs
    w = meshmod.face_area_pair_weights(m)
    g1 = orc.Gamg(a, w, 10)
    _, p_direct, _ = g1.solve(M, "GaussSeidel", np.zeros(m.nCells), b, tolerance=1e-8, maxIter=100)
    # ICCG / BICCG on the coarsest level (GAMGSolverSolve.C:568-606) instead of the LU solve
    psi, perf, hist = g1.solve(M, "GaussSeidel", np.zeros(m.nCells), b, tolerance=1e-8, maxIter=100,
                               directSolveCoarsest=0)
    assert perf.converged and abs(perf.nIterations - p_direct.nIterations) <= 2
    np.testing.assert_allclose(psi, xs, atol=1e-5)
    g2 = orc.Gamg(a, w, 10, mergeLevels=2)
    psi, perf, hist = g2.solve(M, "GaussSeidel", np.zeros(m.nCells), b, tolerance=1e-8, maxIter=200)
    assert perf.converged and np.all(np.diff(hist) < 0)
    np.testing.assert_allclose(psi, xs, atol=1e-5)


def test_gamg_no_levels_is_error(meshmod, orc):
    m = meshmod.hex_mesh(2)
    a = orc.Addr(m.nCells, m.lower, m.upper)
    c = meshmod.pressure_laplacian(m)
    M = orc.Matrix(a, c["diag"], c["upper"], None)
    g = orc.Gamg(a, meshmod.face_area_pair_weights(m), nCellsInCoarsestLevel=10)
    assert g.nLevels == 0
    with pytest.raises(RuntimeError):
        g.solve(M, "GaussSeidel", np.zeros(8), np.ones(8))


def test_fv_face_sums_vs_numpy(meshmod, orc):
    m = meshmod.hex_mesh(7, 5, 6)
    a = orc.Addr(m.nCells, m.lower, m.upper)
    rng = np.random.default_rng(0)
    bfc = np.concatenate([p.faceCells for p in m.patches]).astype(np.int32)
    bSf = np.concatenate([p.Sf for p in m.patches])
    V = m.volumes() * rng.uniform(0.9, 1.1, m.nCells)
    Sf = m.Sf()
    for nc in (1, 3):
        ssf = rng.uniform(-1, 1, (m.nFaces, nc))
        bssf = rng.uniform(-1, 1, (len(bfc), nc))
        ref = np.zeros((m.nCells, nc))
        np.add.at(ref, m.lower, ssf)
        np.subtract.at(ref, m.upper, ssf)
        np.add.at(ref, bfc, bssf)
        got = np.asarray(orc.surface_integrate(a, ssf.ravel(), bfc, bssf.ravel(), V, nc)).reshape(m.nCells, nc)
        np.testing.assert_allclose(got, ref / V[:, None], rtol=1e-12, atol=1e-12)
        gref = np.zeros((m.nCells, 3, nc))
        np.add.at(gref, m.lower, Sf[:, :, None] * ssf[:, None, :])
        np.subtract.at(gref, m.upper, Sf[:, :, None] * ssf[:, None, :])
        np.add.at(gref, bfc, bSf[:, :, None] * bssf[:, None, :])
        gg = orc.gauss_grad(a, Sf.ravel(), ssf.ravel(), bfc, bSf.ravel(), bssf.ravel(), V, nc)
        np.testing.assert_allclose(gg.reshape(m.nCells, 3, nc), gref / V[:, None, None], rtol=1e-12, atol=1e-12)
    # Gauss gradient of a linear field is exact on a uniform hex mesh with exact face values
    cc, fc = m.cell_centres(), m.face_centres()
    g = np.array([1.5, -2.0, 0.5])
    bfcen = np.concatenate([cc[p.faceCells] + 0.5 * m.h * np.sign(p.Sf[0]) * (np.abs(p.Sf[0]) > 0) for p in m.patches])
    grad = orc.gauss_grad(a, Sf.ravel(), fc @ g, bfc, bSf.ravel(), bfcen @ g, m.volumes(), 1)
    np.testing.assert_allclose(grad, np.tile(g, (m.nCells, 1)), atol=1e-10)
    # divergence of a constant flux field vanishes in the interior
    up, dg = orc.laplacian_fill(a, m.deltaCoeffs(), m.magSf())
    A = dense_from_ldu(m.nCells, m.lower, m.upper, dg, up)
    np.testing.assert_allclose(A.sum(1), 0, atol=1e-12)
    lo, up2, dg2 = orc.convection_fill(a, m.weights(), rng.uniform(-1, 1, m.nFaces))
    A2 = dense_from_ldu(m.nCells, m.lower, m.upper, dg2, up2, lo)
    np.testing.assert_allclose(A2.sum(0), 0, atol=1e-12)   # conservative: column sums vanish
    vf = rng.uniform(-1, 1, m.nCells)
    np.testing.assert_allclose(orc.interpolate_linear(a, m.weights(), vf), 0.5 * (vf[m.lower] + vf[m.upper]))


@pytest.mark.parametrize("kind", ["P", "U"])
def test_gamg_over_cyclic_interfaces(meshmod, orc, kind):
    """cyclicGAMGInterface (cyclicGAMGInterface.C:70-150): the coarse faces of a cyclic pair are the unique (owner-side
    coarse cell, other-side coarse cell) pairs in order of appearance -- both patches of the pair enumerate the same
    faces, face i of one facing face i of the other; the V-cycle converges to the periodic solution."""
    from test_oracle_core import _cyclic_case
    m, c, ps, fc, nr, lo, hi = _cyclic_case(meshmod, kind)
    a = orc.Addr(m.nCells, m.lower, m.upper, ps, fc, neighbRank=nr)
    M = orc.Matrix(a, c["diag"], c["upper"], c["lower"], c["bou"], c["int"])
    g = orc.Gamg(a, meshmod.face_area_pair_weights(m), 4, 1)
    assert g.nLevels >= 2
    for lev in range(g.nLevels):
        la = g.level_addr(lev)
        pstart, cells = la.patch_start(), la.face_cells()
        n0, n1 = pstart[1] - pstart[0], pstart[2] - pstart[1]
        assert n0 == n1 > 0                                   # the two sides hold the same coarse faces
        r = g.restrict_addr(lev)
        pfr = g.patch_face_restrict(lev)
        fine = a if lev == 0 else g.level_addr(lev - 1)
        fps, ffc = fine.patch_start(), fine.face_cells()
        nf = fps[1] - fps[0]
        # fine face i of patch 0 and fine face i of patch 1 land on coarse faces with the same index in their patches
        assert np.array_equal(pfr[:nf] - pstart[0], pfr[nf:2 * nf] - pstart[1])
        # and a coarse patch face sits on the image of its fine faces' cells
        assert np.array_equal(cells[pfr[: 2 * nf]], r[ffc[: 2 * nf]])
    x = np.random.default_rng(2).standard_normal(m.nCells)
    b = M.amul(x)
    psi, perf, hist = g.solve(M, "GaussSeidel", np.zeros(m.nCells), b, tolerance=1e-10, maxIter=300)
    assert perf.converged
    np.testing.assert_allclose(psi, x, rtol=0, atol=1e-7)
