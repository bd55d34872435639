"""GAMG on the GPU vs the oracle: identical agglomeration (host algorithm on both sides,
compared level by level), residual history per V-cycle within rel 1e-8, identical cycle
counts."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    capi = importlib.import_module("rapidcfd-dev_b200.capi")
    ctx = capi.Context(0)
    yield capi, ctx, torch
    ctx.close()


def _setup(gpu, meshmod, orc, dims, kind, centres=True, merge=1):
    capi, ctx, torch = gpu
    m = meshmod.hex_mesh(*dims)
    c = meshmod.pressure_laplacian(m) if kind == "P" else meshmod.momentum_matrix(m)
    oa = orc.Addr(m.nCells, m.lower, m.upper)
    om = orc.Matrix(oa, c["diag"], c["upper"], c["lower"])
    w = meshmod.face_area_pair_weights(m)
    og = orc.Gamg(oa, w, 10, mergeLevels=merge)
    addr = capi.mesh_to_device(ctx, m, with_centres=centres)
    mat = capi.LduMatrix(addr)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    d = {k: (t(v) if v is not None and len(v) else None) for k, v in c.items()}
    mat.set(d["diag"], d["upper"], d["lower"])
    gg = capi.GamgAgglomeration(addr, w, 10, mergeLevels=merge)
    return m, c, om, og, addr, mat, gg, t, d


@pytest.mark.parametrize("dims,centres,merge", [((12, 10, 8), True, 1), ((16, 16, 16), False, 1),
                                                ((12, 10, 8), True, 2), ((16, 16, 16), True, 3)])
def test_agglomeration_identical(gpu, meshmod, orc, dims, centres, merge):
    m, c, om, og, addr, mat, gg, t, d = _setup(gpu, meshmod, orc, dims, "P", centres, merge)
    assert gg.nLevels == og.nLevels and gg.forward == og.forward
    for lev in range(gg.nLevels):
        assert gg.level_size(lev) == (og.ncells(lev), og.nfaces(lev))
        assert np.array_equal(gg.restrict_addr(lev), og.restrict_addr(lev))
    gg.close()
    mat.close()
    addr.close()


@pytest.mark.parametrize("kind,kw", [("P", {}), ("P", dict(nPreSweeps=1)), ("P", dict(interpolateCorrection=1)),
                                     ("U", {}), ("U", dict(scaleCorrection=1, nPreSweeps=2, nFinestSweeps=1)),
                                     ("P", dict(nFinestSweeps=3, nPostSweeps=1, maxPostSweeps=2))])
def test_gamg_history(gpu, meshmod, orc, kind, kw):
    capi, ctx, torch = gpu
    m, c, om, og, addr, mat, gg, t, d = _setup(gpu, meshmod, orc, (16, 14, 12), kind)
    xs = meshmod.cell_field_global(m, 42)
    b = om.amul(xs)
    ctl = dict(tolerance=1e-8, maxIter=100, **kw)
    psi_ref, pr, href = og.solve(om, "GaussSeidel", np.zeros(m.nCells), b, **ctl)
    psi = torch.zeros(m.nCells, dtype=torch.float64, device=ctx.device)
    perf, hist = mat.solve("GAMG", "GaussSeidel", psi, t(b), gamg=gg, histCap=256, **ctl)
    assert perf.solverName == b"GAMG"
    assert perf.nIterations == pr.nIterations, (perf.nIterations, pr.nIterations)
    assert len(hist) == len(href)
    np.testing.assert_allclose(hist, href, rtol=1e-8)
    np.testing.assert_allclose(psi.cpu().numpy(), psi_ref, rtol=0, atol=1e-9)
    np.testing.assert_allclose(psi.cpu().numpy(), xs, rtol=0, atol=1e-5)
    # second solve on the same handles (cached agglomeration, refreshed coarse matrices)
    psi2 = torch.zeros(m.nCells, dtype=torch.float64, device=ctx.device)
    perf2, hist2 = mat.solve("GAMG", "GaussSeidel", psi2, t(b), gamg=gg, histCap=256, **ctl)
    assert np.array_equal(hist2, hist) and torch.equal(psi2, psi)
    gg.close()
    mat.close()
    addr.close()


@pytest.mark.parametrize("kind,merge", [("P", 2), ("U", 2), ("U", 3)])
def test_gamg_merge_levels_history(gpu, meshmod, orc, kind, merge):
    """mergeLevels > 1 (combineLevels): composed maps, including the reference's flip rule for the
    asymmetric coarse matrices, give the oracle's cycle history."""
    capi, ctx, torch = gpu
    m, c, om, og, addr, mat, gg, t, d = _setup(gpu, meshmod, orc, (16, 14, 12), kind, merge=merge)
    xs = meshmod.cell_field_global(m, 42)
    b = om.amul(xs)
    ctl = dict(tolerance=1e-8, maxIter=200)
    psi_ref, pr, href = og.solve(om, "GaussSeidel", np.zeros(m.nCells), b, **ctl)
    psi = torch.zeros(m.nCells, dtype=torch.float64, device=ctx.device)
    perf, hist = mat.solve("GAMG", "GaussSeidel", psi, t(b), gamg=gg, histCap=256, **ctl)
    assert perf.nIterations == pr.nIterations, (perf.nIterations, pr.nIterations)
    np.testing.assert_allclose(hist, href, rtol=1e-8)
    np.testing.assert_allclose(psi.cpu().numpy(), psi_ref, rtol=0, atol=1e-9)
    gg.close()
    mat.close()
    addr.close()


@pytest.mark.parametrize("kind", ["P", "U"])
def test_gamg_krylov_coarsest(gpu, meshmod, orc, kind):
    """directSolveCoarsest false: ICCG (symmetric) / BICCG (asymmetric) on the coarsest level to
    the GAMG tolerances (GAMGSolverSolve.C:568-606).  The nested solve stops on a tolerance, so
    the outer history is compared to 1e-5 rather than to rounding."""
    capi, ctx, torch = gpu
    m, c, om, og, addr, mat, gg, t, d = _setup(gpu, meshmod, orc, (16, 14, 12), kind)
    xs = meshmod.cell_field_global(m, 42)
    b = om.amul(xs)
    ctl = dict(tolerance=1e-8, maxIter=100, directSolveCoarsest=0)
    psi_ref, pr, href = og.solve(om, "GaussSeidel", np.zeros(m.nCells), b, **ctl)
    psi = torch.zeros(m.nCells, dtype=torch.float64, device=ctx.device)
    perf, hist = mat.solve("GAMG", "GaussSeidel", psi, t(b), gamg=gg, histCap=256, **ctl)
    assert perf.converged and perf.nIterations == pr.nIterations, (perf.nIterations, pr.nIterations)
    np.testing.assert_allclose(hist, href, rtol=1e-5)
    np.testing.assert_allclose(psi.cpu().numpy(), xs, rtol=0, atol=1e-5)
    # and the default (LU) coarsest solve still works on the same handles afterwards
    psi2 = torch.zeros(m.nCells, dtype=torch.float64, device=ctx.device)
    perf2, _ = mat.solve("GAMG", "GaussSeidel", psi2, t(b), gamg=gg, tolerance=1e-8, maxIter=100)
    assert perf2.converged
    gg.close()
    mat.close()
    addr.close()


def test_gamg_loop_semantics_and_errors(gpu, meshmod, orc):
    capi, ctx, torch = gpu
    m, c, om, og, addr, mat, gg, t, d = _setup(gpu, meshmod, orc, (10, 10, 10), "P")
    b = meshmod.cell_field_global(m, 7)
    for ctl in (dict(tolerance=0.0, maxIter=3), dict(tolerance=1e30, maxIter=9, minIter=2),
                dict(tolerance=1e30, maxIter=9), dict(tolerance=0.0, relTol=0.01, maxIter=50)):
        _, pr, href = og.solve(om, "GaussSeidel", np.zeros(m.nCells), b, **ctl)
        psi = torch.zeros(m.nCells, dtype=torch.float64, device=ctx.device)
        perf, hist = mat.solve("GAMG", "GaussSeidel", psi, t(b), gamg=gg, histCap=64, **ctl)
        assert perf.nIterations == pr.nIterations, ctl
        np.testing.assert_allclose(hist, href, rtol=1e-8)
    with pytest.raises(capi.B200LduError):
        mat.solve("GAMG", "GaussSeidel", torch.zeros(m.nCells, dtype=torch.float64, device=ctx.device), t(b))
    tiny = meshmod.hex_mesh(2)
    ta = capi.mesh_to_device(ctx, tiny)
    tg = capi.GamgAgglomeration(ta, meshmod.face_area_pair_weights(tiny), 10)
    assert tg.nLevels == 0
    tm = capi.LduMatrix(ta)
    cc = meshmod.pressure_laplacian(tiny)
    tm.set(t(cc["diag"]), t(cc["upper"]))
    with pytest.raises(capi.B200LduError) as e:
        tm.solve("GAMG", "GaussSeidel", torch.zeros(8, dtype=torch.float64, device=ctx.device),
                 torch.ones(8, dtype=torch.float64, device=ctx.device), gamg=tg)
    assert e.value.rc == -8
    for h in (tg, tm, ta, gg, mat, addr):
        h.close()


def test_gamg_large(gpu, meshmod, orc):
    """64^3: many levels, 512-row bands; converges like the oracle."""
    capi, ctx, torch = gpu
    m, c, om, og, addr, mat, gg, t, d = _setup(gpu, meshmod, orc, (64, 64, 64), "P")
    b = meshmod.cell_field_global(m, 9)
    ctl = dict(tolerance=1e-6, maxIter=60)
    _, pr, href = og.solve(om, "GaussSeidel", np.zeros(m.nCells), b, **ctl)
    psi = torch.zeros(m.nCells, dtype=torch.float64, device=ctx.device)
    perf, hist = mat.solve("GAMG", "GaussSeidel", psi, t(b), gamg=gg, histCap=128, **ctl)
    assert gg.nLevels == og.nLevels >= 10
    assert abs(perf.nIterations - pr.nIterations) <= 1
    k = min(len(hist), len(href), 12)
    np.testing.assert_allclose(hist[:k], href[:k], rtol=1e-7)
    assert perf.converged
    for h in (gg, mat, addr):
        h.close()
