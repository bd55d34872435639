"""GPU diagnostic: per-iteration relative deviation of the solver residual histories from the oracle's
(decides the tolerances stated in tests/test_gpu_parity.py).  python tools/diag_hist.py > gpurun_out/diag_hist.txt"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import ldu_oracle as orc  # noqa: E402

capi = importlib.import_module("rapidcfd-dev_b200.capi")
meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
ctx = capi.Context(0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
for dims in ((16, 12, 10), (32, 32, 32), (64, 64, 64)):
    for kind, solvers in (("P", (("PCG", "DIC"), ("PCG", "diagonal"))),
                          ("U", (("PBiCG", "DILU"), ("PBiCG", "none"), ("PBiCGStab", "DILU"), ("PBiCGStab", "diagonal")))):
        m = meshmod.hex_mesh(*dims)
        c = meshmod.pressure_laplacian(m) if kind == "P" else meshmod.momentum_matrix(m)
        oa = orc.Addr(m.nCells, m.lower, m.upper)
        om = orc.Matrix(oa, c["diag"], c["upper"], c["lower"])
        addr = capi.mesh_to_device(ctx, m)
        mat = capi.LduMatrix(addr)
        mat.set(t(c["diag"]), t(c["upper"]), t(c["lower"]) if c["lower"] is not None else None)
        b = om.amul(meshmod.cell_field_global(m, 42))
        for solver, pre in solvers:
            kw = dict(tolerance=1e-10, maxIter=60)
            _, pr, href = om.solve(solver, pre, np.zeros(m.nCells), b, **kw)
            psi = torch.zeros(m.nCells, dtype=torch.float64, device=ctx.device)
            perf, hist = mat.solve(solver, pre, psi, t(b), histCap=128, **kw)
            k = min(len(hist), len(href), 40)
            h, hr = np.asarray(hist[:k]), np.asarray(href[:k])
            rel = np.abs(h - hr) / np.abs(hr)
            print(dims, kind, solver, pre, "its", perf.nIterations, pr.nIterations,
                  "rel@10 %.1e @20 %.1e @30 %.1e max40 %.1e" % (rel[:10].max(), rel[:20].max(), rel[:30].max(), rel.max()),
                  "res@30 %.1e" % (hr[min(30, k - 1)]), flush=True)
        mat.close()
        addr.close()
ctx.close()
