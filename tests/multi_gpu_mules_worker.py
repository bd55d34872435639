"""Worker for tests/test_gpu_mules.py::test_multi_gpu_mules (torchrun, backend nccl, one rank per GPU): explicit MULES over
processor patches -- neighbour values in the extrema, the coupled face rule, the minimum with the other side after every
sweep -- against the N-rank ORACLE (all ranks in lockstep in this process) on the same decomposition, bit for bit."""
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
    from oracle import mules_oracle as mo
    import test_mules_cpu as tm
    capi = importlib.import_module("rapidcfd-dev_b200.capi")
    meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
    mules = importlib.import_module("rapidcfd-dev_b200.mules")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = capi.Context(local)
    ctx.comm_init_from_torch()
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(ctx.device)
    for combo in ("one-zero", "rho-SpSu"):
        cases, _, _, exchange, meshes = tm.decomposed(meshmod, (16, 12, 8), world, seed=21, combo=combo)
        want = mo.limiter_ranks(cases, exchange)
        c, m = cases[rank], meshes[rank]
        nC, nB = c["nCoupled"], len(c["bFaceCells"])
        addr = capi.mesh_to_device(ctx, m)
        capi.fv_boundary_set(addr, c["bFaceCells"])
        ops = capi.FieldOps(ctx)
        psi = t(c["psi"])
        pnf = capi.fv_patch_neighbour_field(addr, 1, psi)
        assert np.array_equal(pnf.cpu().numpy(), c["psiB"][nB - nC:]), f"rank {rank}: patchNeighbourField differs"
        psiB = torch.cat([t(c["psiB"][: nB - nC]), pnf])
        kw = {k: t(c.get(k)) for k in ("rho", "rho0", "Sp", "Su")}
        lam, lamB = capi.mules_limiter(addr, t(c["V"]), c["rDeltaT"], psi, t(c["psi0"]), psiB, t(c["phiBD"]), t(c["phiBDB"]),
                                       t(c["phiCorr"]), t(c["phiCorrB"]), 1.0, 0.0, 3, kw["rho"], kw["rho0"], kw["Sp"], kw["Su"], nC)
        assert np.array_equal(lam.cpu().numpy(), want[rank][0]), f"rank {rank}: lambda differs"
        assert np.array_equal(lamB.cpu().numpy(), want[rank][1]), f"rank {rank}: boundary lambda differs"
        print("MULTI-GPU-MULES-OK limiter", combo, flush=True)
        # limit + explicitSolve through the compositions: phi / phiPsi rebuilt from the case's pieces
        phiPsi, phiPsiB = c["phiBD"] + c["phiCorr"], c["phiBDB"] + c["phiCorrB"]
        phi, phiB = tm.decomposed_fluxes(meshmod, (16, 12, 8), world, rank, seed=21)
        cfc = torch.from_numpy(np.ascontiguousarray(c["bFaceCells"][nB - nC:], np.int32)).to(ctx.device)
        lp, lpB = mules.limit(capi, addr, ops, t(c["V"]), c["rDeltaT"], psi, t(c["psi0"]), psiB, t(phi), t(phiB), t(phiPsi), t(phiPsiB),
                              1.0, 0.0, 3, kw["rho"], kw["rho0"], kw["Sp"], kw["Su"], nC, cfc)
        wl, wlB = c["phiBD"] + want[rank][0] * c["phiCorr"], c["phiBDB"] + want[rank][1] * c["phiCorrB"]
        assert np.array_equal(lp.cpu().numpy(), wl) and np.array_equal(lpB.cpu().numpy(), wlB), f"rank {rank}: limited flux differs"
        new = mules.explicit_solve(capi, addr, ops, t(c["V"]), c["rDeltaT"], t(c["psi0"]), lp, lpB, kw["rho"], kw["rho0"], kw["Sp"], kw["Su"])
        ref = mo.explicit_solve(c["nCells"], c["lower"], c["upper"], c["bFaceCells"], c["V"], c["rDeltaT"], c["psi0"], wl, wlB,
                                c.get("rho"), c.get("rho0"), c.get("Sp"), c.get("Su"))
        assert np.array_equal(new.cpu().numpy(), ref), f"rank {rank}: explicitSolve differs"
        # conservation across the processor faces: what leaves one side enters the other
        sent = exchange_torch(dist, lpB[nB - nC:], m, rank)
        assert torch.equal(sent, -lpB[nB - nC:]), f"rank {rank}: limited fluxes of a processor face are not opposite"
        print("MULTI-GPU-MULES-OK update", combo, flush=True)
        addr.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


def exchange_torch(dist, mine, m, rank):
    """the other side's values of this rank's processor faces (torch.distributed point to point, patch by patch)"""
    import torch
    out = torch.empty_like(mine)
    ops_, o = [], 0
    for p in m.coupled_patches():
        k = len(p.faceCells)
        ops_.append(dist.P2POp(dist.isend, mine[o: o + k].contiguous(), p.neighbRank))
        ops_.append(dist.P2POp(dist.irecv, out[o: o + k], p.neighbRank))
        o += k
    for w in dist.batch_isend_irecv(ops_):
        w.wait()
    torch.cuda.synchronize()
    return out


if __name__ == "__main__":
    main()
