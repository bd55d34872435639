"""Worker for test_gpu_multi.py (torchrun, backend nccl, one rank per GPU): the CUDA path
over the brick decomposition (NCCL halo exchange + device all-reduce) against the
N-rank ORACLE run (in-process threads) on the same decomposition."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import dist_helpers as dh
    from oracle import ldu_oracle as orc
    capi = importlib.import_module("rapidcfd-dev_b200.capi")
    meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 16
    ctx = capi.Context(local)
    ctx.comm_init_from_torch()
    dev = ctx.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for kind, solver, pre in (("P", "PCG", "DIC"), ("P", "PCG", "diagonal"), ("U", "PBiCG", "DILU"),
                              ("U", "PBiCGStab", "none"), ("U", "smoothSolver", "GaussSeidel")):
        gm, gc = dh.global_case(meshmod, n, kind)
        ga, gM = dh.oracle_matrix(orc, gm, gc)
        xs = meshmod.cell_field_global(gm, 42)
        b = gM.amul(xs)
        x = meshmod.cell_field_global(gm, 3)
        kw = dict(tolerance=1e-8, maxIter=300)
        # N-rank oracle (threads) -- the reference for the decomposed run
        ex = dh.ThreadExchange(world)

        def rank_fn(r):
            m, c = dh.local_case(meshmod, n, world, r, kind)
            a, M = dh.oracle_matrix(orc, m, c)
            comm = ex.comm(orc, r, m, n ** 3)
            y = M.amul(x[m.cellGlobal], comm)
            yT = M.tmul(x[m.cellGlobal], comm)
            psi, perf, hist = M.solve(solver, pre, np.zeros(m.nCells), b[m.cellGlobal], comm=comm, **kw)
            return y, yT, psi, perf.nIterations, hist
        oy, oyT, opsi, onit, ohist = dh.run_threads(world, rank_fn)[rank]

        mesh, coef = dh.local_case(meshmod, n, world, rank, kind)
        addr = capi.mesh_to_device(ctx, mesh)
        mat = capi.LduMatrix(addr)
        d = {k: (t(v) if v is not None and len(v) else None) for k, v in coef.items()}
        mat.set(d["diag"], d["upper"], d["lower"], d["bou"], d["int"])
        y = mat.Amul(t(x[mesh.cellGlobal])).cpu().numpy()
        assert np.array_equal(y, oy), f"rank {rank}: Amul with halo differs ({kind})"
        yT = mat.Tmul(t(x[mesh.cellGlobal])).cpu().numpy()
        assert np.array_equal(yT, oyT), f"rank {rank}: Tmul with halo differs ({kind})"
        psi = torch.zeros(mesh.nCells, dtype=torch.float64, device=dev)
        perf, hist = mat.solve(solver, pre, psi, t(b[mesh.cellGlobal]), histCap=512, **kw)
        assert abs(perf.nIterations - onit) <= 1, (solver, pre, perf.nIterations, onit)
        k = min(15, len(hist), len(ohist))
        assert np.allclose(hist[:k], ohist[:k], rtol=1e-8, atol=0), (solver, pre, hist[:k], ohist[:k])
        assert np.allclose(psi.cpu().numpy(), xs[mesh.cellGlobal], atol=1e-5)
        mat.close()
        addr.close()
        dist.barrier()
        if rank == 0:
            print(f"MULTI-GPU-OK {solver} {pre} ranks={world} iterations={perf.nIterations}", flush=True)
    # ---- GAMG across the ranks: processor-interface agglomeration, restricted interface
    # coefficients, peer-memory gather for the global coarsest solve ----
    for kind, kw, merge, htol in (("P", {}, 1, 1e-7), ("U", dict(nPreSweeps=1), 1, 1e-7),
                                  ("U", {}, 2, 1e-7),                        # combineLevels across processor patches
                                  ("P", dict(directSolveCoarsest=0), 1, 1e-4)):  # ICCG on the coarsest level
        n = 16
        gm, gc = dh.global_case(meshmod, n, kind)
        ga, gM = dh.oracle_matrix(orc, gm, gc)
        xs = meshmod.cell_field_global(gm, 42)
        b = gM.amul(xs)
        ctl = dict(tolerance=1e-8, maxIter=60, **kw)
        ex = dh.ThreadExchange(world)

        def rank_fn(r):
            m, c = dh.local_case(meshmod, n, world, r, kind)
            a, M = dh.oracle_matrix(orc, m, c)
            comm = ex.comm(orc, r, m, n ** 3)
            g = orc.Gamg(a, meshmod.face_area_pair_weights(m), 10, mergeLevels=merge, comm=comm)
            maps = [g.restrict_addr(k) for k in range(g.nLevels)]
            sizes = [(g.ncells(k), g.nfaces(k)) for k in range(g.nLevels)]
            psi, perf, hist = g.solve(M, "GaussSeidel", np.zeros(m.nCells), b[m.cellGlobal], comm=comm, **ctl)
            return maps, sizes, psi, perf.nIterations, hist
        omaps, osizes, opsi, onit, ohist = dh.run_threads(world, rank_fn)[rank]

        mesh, coef = dh.local_case(meshmod, n, world, rank, kind)
        addr = capi.mesh_to_device(ctx, mesh)
        mat = capi.LduMatrix(addr)
        d = {k: (t(v) if v is not None and len(v) else None) for k, v in coef.items()}
        mat.set(d["diag"], d["upper"], d["lower"], d["bou"], d["int"])
        gg = capi.GamgAgglomeration(addr, meshmod.face_area_pair_weights(mesh), 10, mergeLevels=merge)
        assert gg.nLevels == len(omaps), (gg.nLevels, len(omaps))
        for k in range(gg.nLevels):
            assert gg.level_size(k) == osizes[k]
            assert np.array_equal(gg.restrict_addr(k), omaps[k])
        psi = torch.zeros(mesh.nCells, dtype=torch.float64, device=dev)
        perf, hist = mat.solve("GAMG", "GaussSeidel", psi, t(b[mesh.cellGlobal]), gamg=gg, histCap=128, **ctl)
        assert perf.nIterations == onit, (perf.nIterations, onit)
        assert np.allclose(hist, ohist, rtol=htol, atol=0), (hist, ohist)
        assert np.allclose(psi.cpu().numpy(), opsi, atol=1e-8 if htol < 1e-6 else 1e-6)
        assert np.allclose(psi.cpu().numpy(), xs[mesh.cellGlobal], atol=1e-5)
        gg.close()
        mat.close()
        addr.close()
        dist.barrier()
        if rank == 0:
            print(f"MULTI-GPU-OK GAMG {kind} ranks={world} cycles={perf.nIterations} levels={len(omaps)} "
                  f"mergeLevels={merge} {kw}", flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
