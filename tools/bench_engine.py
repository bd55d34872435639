"""Engine micro-benchmark for A/B builds (B200LDU_LIB=...): Amul / AINV+dot / Jacobi / sumA kernels alone and a
50-iteration PCG solve, at --n 256 and 128.  One line per size."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

capi = importlib.import_module("rapidcfd-dev_b200.capi")
meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="256,128")
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
ctx = capi.Context(0)
dev = ctx.device
L = capi.lib()
tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for n in [int(x) for x in args.sizes.split(",")]:
    mesh = meshmod.hex_mesh(n)
    coef = meshmod.pressure_laplacian(mesh)
    addr = capi.mesh_to_device(ctx, mesh)
    mat = capi.LduMatrix(addr)
    mat.set(tt(coef["diag"]), tt(coef["upper"]))
    vl = addr.vec_len
    xb = torch.rand(vl, dtype=torch.float64, device=dev)
    yb = torch.zeros(vl, dtype=torch.float64, device=dev)
    bb = torch.rand(vl, dtype=torch.float64, device=dev)
    out = {"lib": os.path.basename(capi.LIB_PATH), "n": n, "bandRows": addr.info()["bandRows"]}
    for op in ("amul", "ainv_dot", "jacobi", "sumA"):
        f = lambda: capi.check(L.b200ldu_bench_op(mat.h, op.encode(), capi._dp(xb), capi._dp(yb), capi._dp(bb)))
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            f()
        e1.record()
        torch.cuda.synchronize()
        out[op + "_us"] = round(1e3 * e0.elapsed_time(e1) / args.reps, 1)
    b = tt(meshmod.cell_field_global(mesh, 9))
    psi = torch.zeros(mesh.nCells, dtype=torch.float64, device=dev)
    its = 50
    for _ in range(3):
        psi.zero_()
        perf, _ = mat.solve("PCG", "DIC", psi, b, tolerance=0.0, maxIter=its - 1)
    assert perf.nIterations == its
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        psi.zero_()
        mat.solve("PCG", "DIC", psi, b, tolerance=0.0, maxIter=its - 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out["pcg_Mcell_iters_s"] = round(n ** 3 * its / (ms * 1e-3) / 1e6)
    out["pcg_us_per_iter"] = round(1e3 * ms / its, 1)
    print(json.dumps(out), flush=True)
    mat.close()
    addr.close()
ctx.close()
