"""Worker for tests/test_zzz_fvm_gpu.py::test_multi_gpu_fvm (torchrun, backend nccl, one rank per GPU): the fvMatrix
glue over processor patches -- patchNeighbourField exchange, H, A, flux, residual, the component loop of
solveSegregated -- against the N-rank ORACLE run (in-process threads) on the same decomposition."""
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
    from oracle import ldu_oracle as orc
    import test_oracle_fvm as tf
    capi = importlib.import_module("rapidcfd-dev_b200.capi")
    meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = capi.Context(local)
    ctx.comm_init_from_torch()
    t = lambda a: torch.from_numpy(np.array(a, dtype=np.float64).ravel()).to(ctx.device)
    host = lambda x, nc: x.cpu().numpy().reshape(-1, nc)
    ctl = dict(tolerance=1e-12, maxIter=800)
    G = tf.decomposed_global(meshmod, orc, 16)
    R = tf.oracle_rank_results(meshmod, orc, G, world, ctl)[rank]
    m, d, cou = tf.decomposed_rank(meshmod, G, world, rank)
    x = G["x"][m.cellGlobal]
    addr = capi.mesh_to_device(ctx, m)
    capi.fv_boundary_set(addr, d["bfc"])
    mat = capi.LduMatrix(addr)
    diag, upper = t(d["diag"]), t(d["upper"])
    bou = inn = t(-cou)
    mat.set(diag, upper, None, bou, inn)
    nC = len(cou)
    psi = t(x)
    fv = capi.FvMatrix(mat, 3, diag, t(d["source"]), psi, t(d["V"]), t(d["ic"]), t(d["bc"]))
    pnf = capi.fv_patch_neighbour_field(addr, 3, psi)
    assert np.array_equal(host(pnf, 3), R["pnf"]), f"rank {rank}: patchNeighbourField differs"
    assert np.array_equal(host(fv.A(), 1)[:, 0], R["A"]), f"rank {rank}: A differs"
    assert np.array_equal(host(fv.H(pnf), 3), R["H"]), f"rank {rank}: H differs"
    fi, fb, fc = fv.flux(len(d["bfc"]), nC, pnf)
    assert np.array_equal(host(fi, 3), R["internal"]) and np.array_equal(host(fb, 3), R["boundary"])
    assert np.array_equal(host(fc, 3), R["coupled"]), f"rank {rank}: coupled flux differs"
    print("MULTI-GPU-FVM-OK operations", flush=True)
    # scalar residual (component 0)
    psi1 = t(x[:, 0])
    fv1 = capi.FvMatrix(mat, 1, diag, t(d["source"][:, 0]), psi1, t(d["V"]), t(d["ic"][:, 0]), t(d["bc"][:, 0]))
    res = fv1.residual(capi.fv_patch_neighbour_field(addr, 1, psi1))
    np.testing.assert_allclose(res.cpu().numpy(), R["res"], rtol=1e-12, atol=1e-13)
    print("MULTI-GPU-FVM-OK residual", flush=True)
    # component loop: the coupled boundary source goes in with the start field's neighbour values, and out again
    psi0 = t(np.zeros_like(x))
    fz = capi.FvMatrix(mat, 3, diag, t(d["source"]), psi0, t(d["V"]), t(d["ic"]), t(d["bc"]))
    perfs = fz.solve("PCG", "DIC", pnf=capi.fv_patch_neighbour_field(addr, 3, psi0), **ctl)
    assert all(p.converged for p in perfs)
    assert all(abs(p.nIterations - k) <= 2 for p, k in zip(perfs, R["nIter"])), (rank, [p.nIterations for p in perfs], R["nIter"])
    np.testing.assert_allclose(host(psi0, 3), R["psi"], rtol=0, atol=1e-8)
    print("MULTI-GPU-FVM-OK solveSegregated", flush=True)
    mat.close()
    addr.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
