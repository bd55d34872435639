"""Worker for tests/test_zzz_fvm_gpu.py::test_multi_gpu_icofoam (torchrun, backend nccl, one rank per GPU): two icoFoam
steps of the brick-decomposed cavity on the devices against the single-domain ORACLE run."""
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
    from oracle import piso_oracle as po
    capi = importlib.import_module("rapidcfd-dev_b200.capi")
    meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
    ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = capi.Context(local)
    ctx.comm_init_from_torch()
    n = 16
    ctl = dict(tolerance=1e-12, relTol=0.0)
    _, ref = po.cavity_from_hex(orc, meshmod, n)
    for _ in range(2):
        ref.step(UControls=ctl, pControls=ctl)

    def allsum(v):
        t = torch.from_numpy(np.array(v, dtype=np.float64)).to(ctx.device)
        dist.all_reduce(t)
        return t.cpu().numpy()
    m, case = ico.cavity_rank(capi, ctx, torch, n, world, rank, allsum)
    for _ in range(2):
        perfs, cont = case.step(UControls=ctl, pControls=ctl)
    assert all(p.converged for p in perfs["U"] + perfs["p"]), f"rank {rank}: a solve did not converge"
    cg = m.cellGlobal
    np.testing.assert_allclose(case.U.cpu().numpy().reshape(-1, 3), ref.U[cg], rtol=0, atol=1e-8)
    np.testing.assert_allclose(case.p.cpu().numpy(), ref.p[cg], rtol=0, atol=1e-8)
    assert cont[-1][0] < 1e-10, cont
    print("MULTI-GPU-ICOFOAM-OK", flush=True)
    case.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
