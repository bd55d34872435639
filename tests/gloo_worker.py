"""Worker for test_distributed_cpu.py::test_gloo_two_process_pcg (launched by torchrun,
backend gloo, world_size 2): oracle PCG over the brick decomposition with halo exchange
and rank-ordered all-reduce through torch.distributed; rank 0 compares with the
single-domain solve."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch.distributed as dist
    import dist_helpers as dh
    from oracle import ldu_oracle as orc
    meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 10
    m, c = dh.local_case(meshmod, n, world, rank, "P")
    a, M = dh.oracle_matrix(orc, m, c)
    comm = dh.torch_comm(orc, m, n ** 3)
    gm, gc = dh.global_case(meshmod, n, "P")
    ga, gM = dh.oracle_matrix(orc, gm, gc)
    xs = meshmod.cell_field_global(gm, 42)
    b = gM.amul(xs)
    kw = dict(tolerance=1e-8, maxIter=400)
    psi, perf, hist = M.solve("PCG", "diagonal", np.zeros(m.nCells), b[m.cellGlobal], comm=comm, **kw)
    psi_ref, pr, href = gM.solve("PCG", "diagonal", np.zeros(gm.nCells), b, **kw)
    assert abs(perf.nIterations - pr.nIterations) <= 1, (perf.nIterations, pr.nIterations)
    k = min(20, len(hist), len(href))
    assert np.allclose(hist[:k], href[:k], rtol=1e-8, atol=0)
    assert np.allclose(psi, psi_ref[m.cellGlobal], atol=1e-6)
    dist.barrier()
    if rank == 0:
        print("GLOO-PCG-OK", perf.nIterations)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
