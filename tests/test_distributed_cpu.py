"""N > 1 host logic on CPU: the brick decomposition + processor-patch coefficients +
halo/all-reduce protocol reproduce the single-domain result (oracle numerics on every
rank).  In-process threads cover 2/4/8 ranks; a real 2-process gloo run covers the
torch.distributed plumbing bench.py uses."""
import os
import subprocess
import sys

import numpy as np
import pytest

import dist_helpers as dh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_decomposition_covers_mesh(meshmod):
    n = 8
    for nR in (2, 4, 8):
        seen = np.zeros(n ** 3, dtype=int)
        nif = 0
        for r in range(nR):
            m = meshmod.decompose(n, nR, r)
            seen[m.cellGlobal] += 1
            for p in m.coupled_patches():
                assert p.kind == "processor" and p.neighbRank != r
                nb = meshmod.decompose(n, nR, p.neighbRank)
                q = [x for x in nb.coupled_patches() if x.neighbRank == r]
                assert len(q) == 1
                # both sides enumerate the patch identically
                assert np.array_equal(p.nbrGlobalCells, nb.cellGlobal[q[0].faceCells])
                assert np.array_equal(q[0].nbrGlobalCells, m.cellGlobal[p.faceCells])
                nif += len(p.faceCells)
            ranks = [p.neighbRank for p in m.coupled_patches()]
            assert ranks == sorted(ranks)
        assert np.all(seen == 1)
        g = meshmod.hex_mesh(n)
        tot_internal = sum(meshmod.decompose(n, nR, r).nFaces for r in range(nR))
        assert tot_internal + nif // 2 == g.nFaces


@pytest.mark.parametrize("nR", [2, 4, 8])
@pytest.mark.parametrize("kind", ["P", "U"])
def test_amul_matches_single_domain(meshmod, orc, nR, kind):
    n = 8
    gm, gc = dh.global_case(meshmod, n, kind)
    ga, gM = dh.oracle_matrix(orc, gm, gc)
    x = meshmod.cell_field_global(gm, 3)
    ref = gM.amul(x)
    refT = gM.tmul(x)
    refS = gM.sumA()
    ex = dh.ThreadExchange(nR)

    def rank_fn(r):
        m, c = dh.local_case(meshmod, n, nR, r, kind)
        a, M = dh.oracle_matrix(orc, m, c)
        comm = ex.comm(orc, r, m, n ** 3)
        xl = x[m.cellGlobal]
        return m.cellGlobal, M.amul(xl, comm), M.tmul(xl, comm), M.sumA()
    for cg, y, yT, s in dh.run_threads(nR, rank_fn):
        np.testing.assert_allclose(y, ref[cg], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(yT, refT[cg], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(s, refS[cg], rtol=1e-13, atol=1e-14)


@pytest.mark.parametrize("nR,solver,pre,kind", [(2, "PCG", "none", "P"), (4, "PCG", "diagonal", "P"),
                                                 (8, "PCG", "DIC", "P"), (2, "PBiCG", "diagonal", "U"),
                                                 (2, "PBiCG", "DILU", "U"), (4, "PBiCGStab", "none", "U"),
                                                 (2, "smoothSolver", "GaussSeidel", "U")])
def test_solver_history_matches_single_domain(meshmod, orc, nR, solver, pre, kind):
    """none/diagonal preconditioning and Jacobi smoothing commute with the decomposition,
    so the decomposed run follows the single-domain residual history up to summation
    order.  AINV ("DIC"/"DILU") has no interface contribution (AINVPreconditioner.C:64-72
    uses upper/lower only): per-rank it is block-local, so only convergence to the same
    solution is required."""
    n = 8
    gm, gc = dh.global_case(meshmod, n, kind)
    ga, gM = dh.oracle_matrix(orc, gm, gc)
    xs = meshmod.cell_field_global(gm, 42)
    b = gM.amul(xs)
    kw = dict(tolerance=1e-8, maxIter=400)
    psi_ref, pr, href = gM.solve(solver, pre, np.zeros(gm.nCells), b, **kw)
    ex = dh.ThreadExchange(nR)

    def rank_fn(r):
        m, c = dh.local_case(meshmod, n, nR, r, kind)
        a, M = dh.oracle_matrix(orc, m, c)
        comm = ex.comm(orc, r, m, n ** 3)
        psi, perf, hist = M.solve(solver, pre, np.zeros(m.nCells), b[m.cellGlobal], comm=comm, **kw)
        return m.cellGlobal, psi, perf.nIterations, hist
    res = dh.run_threads(nR, rank_fn)
    local_precond = pre in ("DIC", "DILU")
    for cg, psi, nit, hist in res:
        if not local_precond:
            assert abs(nit - pr.nIterations) <= 1
            k = min(20, len(hist), len(href))
            np.testing.assert_allclose(hist[:k], href[:k], rtol=1e-8)
        else:
            assert abs(hist[0] - href[0]) <= 1e-12 * href[0]
        np.testing.assert_allclose(psi, psi_ref[cg], atol=1e-6)
    assert len({r[2] for r in res}) == 1  # every rank stops at the same iteration


def test_gloo_two_process_pcg():
    script = os.path.join(ROOT, "tests", "gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", script]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert "GLOO-PCG-OK" in p.stdout


@pytest.mark.parametrize("nR,kind,opts", [(2, "P", {}), (4, "P", {}), (8, "P", {}), (2, "U", {}),
                                          (2, "P", dict(mergeLevels=2)), (4, "U", dict(mergeLevels=2)),
                                          (2, "P", dict(directSolveCoarsest=0))])
def test_gamg_multi_rank_oracle(meshmod, orc, nR, kind, opts):
    """Multi-rank GAMG (processor-interface agglomeration, restricted interface coefficients,
    global coarsest solve): same level count on every rank, matching coarse patch sizes on the
    two sides of every processor patch, monotone convergence to the single-domain solution."""
    n = 12
    gm, gc = dh.global_case(meshmod, n, kind)
    ga, gM = dh.oracle_matrix(orc, gm, gc)
    xs = meshmod.cell_field_global(gm, 42)
    b = gM.amul(xs)
    ex = dh.ThreadExchange(nR)

    def rank_fn(r):
        m, c = dh.local_case(meshmod, n, nR, r, kind)
        a, M = dh.oracle_matrix(orc, m, c)
        comm = ex.comm(orc, r, m, n ** 3)
        g = orc.Gamg(a, meshmod.face_area_pair_weights(m), 10, mergeLevels=opts.get("mergeLevels", 1), comm=comm)
        sizes = [(g.ncells(k), g.npatchfaces(k)) for k in range(g.nLevels)]
        kw = {k: v for k, v in opts.items() if k != "mergeLevels"}
        psi, perf, hist = g.solve(M, "GaussSeidel", np.zeros(m.nCells), b[m.cellGlobal], comm=comm,
                                  tolerance=1e-8, maxIter=100, **kw)
        return m.cellGlobal, psi, perf.nIterations, hist, g.nLevels, sizes, perf.converged
    res = dh.run_threads(nR, rank_fn)
    assert len({r[4] for r in res}) == 1 and res[0][4] >= 2      # same number of levels everywhere
    assert len({r[2] for r in res}) == 1                         # same number of cycles
    for cg, psi, nit, hist, nl, sizes, conv in res:
        assert conv and nit < 60
        assert np.all(np.diff(hist) < 0)
        np.testing.assert_allclose(psi, xs[cg], atol=1e-5)
        assert all(nc >= 10 for nc, _ in sizes)
    # coarse patch faces are created pairwise: totals over the ranks are even on every level
    for k in range(res[0][4]):
        assert sum(r[5][k][1] for r in res) % 2 == 0
    np.testing.assert_allclose(res[0][3], res[-1][3], rtol=0, atol=0)  # identical history on all ranks
