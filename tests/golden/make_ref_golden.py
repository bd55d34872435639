"""Generates tests/golden/ref_golden.npz -- fixtures produced by EXECUTING THE REFERENCE'S OWN SOURCE
(oracle/_ref/libref_*.so: lduMatrixATmul.C, lduMatrixTemplates.C, AINVPreconditioner, JacobiSmoother,
smoothSolver, PCG/PBiCG/PBiCGStab, fvcSurfaceIntegrate.C, gaussGrad.C, pairGAMGAgglomerate.C, GAMGAgglomerateLduAddressing.C, the
agglomeration functors, GAMGSolverSolve/Scale -- compiled for the host, see oracle/ref_harness/).
The oracle is NOT used here: inputs come from rapidcfd-dev_b200/mesh.py and the seeds below, outputs
from oracle/ref_ldu.py only.  Needs /root/reference (to build oracle/_ref); the stored vectors let the
comparison run where the reference tree is absent (GPU box, later rounds).

    python tests/golden/make_ref_golden.py      # rewrites ref_golden.npz
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

OPS_CASES = [("opsP", (13, 7, 5), "P"), ("opsU", (13, 7, 5), "U")]
SOLVE_CASES = [  # (name, dims, kind, solver, preconditioner / smoother, controls)
    ("pcgDIC", (12, 10, 8), "P", "PCG", "DIC", dict(tolerance=1e-7, maxIter=400)),
    ("pcgDiag", (12, 10, 8), "P", "PCG", "diagonal", dict(tolerance=1e-7, maxIter=400)),
    ("pbicgDILU", (12, 10, 8), "U", "PBiCG", "DILU", dict(tolerance=1e-8, maxIter=300)),
    ("pbicgstabDILU", (12, 10, 8), "U", "PBiCGStab", "DILU", dict(tolerance=1e-8, maxIter=300)),
    ("smooth2", (10, 10, 10), "U", "smoothSolver", "GaussSeidel", dict(tolerance=1e-6, maxIter=500, nSweeps=2)),
]
GAMG_CASES = [("gamgP", (16, 14, 12), "P", dict(tolerance=1e-8, maxIter=100)),
              ("gamgU", (16, 14, 12), "U", dict(tolerance=1e-8, maxIter=100)),
              ("gamgPpre", (12, 10, 8), "P", dict(tolerance=1e-8, maxIter=100, nPreSweeps=1, nFinestSweeps=1))]
FV_CASES = [("fv", (9, 7, 5))]   # fvc::surfaceIntegrate, fvc::surfaceSum, gaussGrad<scalar>::gradf
HIST_K = 12   # residual after k loop bodies, k = 1..HIST_K (one reference solve per k, tolerance 0)


def coefficients(meshmod, dims, kind):
    m = meshmod.hex_mesh(*dims)
    c = meshmod.pressure_laplacian(m) if kind == "P" else meshmod.momentum_matrix(m)
    return m, c


def ref_matrix(ref, m, c):
    os_, ls, lo = ref.ldu_arrays(m.nCells, m.lower, m.upper)
    return ref.RefMatrix(m.nCells, m.lower, m.upper, os_, ls, lo, c["diag"], c["upper"], c["lower"]), (os_, ls, lo)


def fv_inputs(meshmod, dims):
    m = meshmod.hex_mesh(*dims)
    rng = np.random.default_rng(11)
    bfc = np.concatenate([p.faceCells for p in m.patches]).astype(np.int32)
    Sf = m.Sf() + rng.uniform(-0.1, 0.1, (m.nFaces, 3))
    bSf = np.concatenate([p.Sf for p in m.patches]) + rng.uniform(-0.1, 0.1, (len(bfc), 3))
    return m, dict(bfc=bfc, Sf=Sf, bSf=bSf, V=m.volumes() * rng.uniform(0.9, 1.1, m.nCells),
                   ssf=rng.uniform(-1, 1, m.nFaces), bssf=rng.uniform(-1, 1, len(bfc)))


def generate(meshmod, ref):
    out = {}
    for name, dims in FV_CASES:
        m, d = fv_inputs(meshmod, dims)
        out[f"{name}.integrate"] = ref.surface_integrate(m.nCells, m.lower, m.upper, d["ssf"], d["bfc"], d["bssf"], d["V"])
        out[f"{name}.sum"] = ref.surface_integrate(m.nCells, m.lower, m.upper, d["ssf"], d["bfc"], d["bssf"], d["V"],
                                                   integrate=False)
        out[f"{name}.grad"] = ref.gauss_gradf(m.nCells, m.lower, m.upper, d["Sf"], d["ssf"], d["bfc"], d["bSf"],
                                              d["bssf"], d["V"])
    for name, dims, kind in OPS_CASES:
        m, c = coefficients(meshmod, dims, kind)
        R, _ = ref_matrix(ref, m, c)
        x, b = meshmod.cell_field_global(m, 3), meshmod.cell_field_global(m, 4)
        for op in ("amul", "tmul", "H", "faceH"):
            out[f"{name}.{op}"] = R.op(op, 0, x)
        out[f"{name}.sumA"] = R.op("sumA", 0)
        out[f"{name}.H1"] = R.op("H1", 0)
        out[f"{name}.residual"] = R.op("residual", 0, x, b)
        for T in (False, True):
            out[f"{name}.ainv.{int(T)}"] = R.ainv(x, False, T)
        out[f"{name}.jacobi1"] = R.jacobi(x, b, 0.9)
    for name, dims, kind, solver, pre, ctl in SOLVE_CASES:
        m, c = coefficients(meshmod, dims, kind)
        R, (os_, ls, lo) = ref_matrix(ref, m, c)
        args = (m.nCells, m.lower, m.upper, os_, ls, lo, c["diag"], c["upper"], c["lower"])
        b = R.op("amul", 0, meshmod.cell_field_global(m, 42))
        z = np.zeros(m.nCells)
        psi, p = ref.solve(solver, pre, *args, z, b, **ctl)
        out[f"{name}.psi"] = psi
        out[f"{name}.perf"] = np.array([p["nIterations"], p["converged"], p["initialResidual"], p["finalResidual"]])
        hist = []
        step = ctl.get("nSweeps", 1) if solver == "smoothSolver" else 1
        for k in range(HIST_K):
            kw = dict(ctl)
            kw.update(tolerance=0.0, maxIter=(k + 1) * step if solver == "smoothSolver" else k)
            hist.append(ref.solve(solver, pre, *args, z, b, **kw)[1]["finalResidual"])
        out[f"{name}.hist"] = np.array(hist)
    for name, dims, kind, ctl in GAMG_CASES:
        m, c = coefficients(meshmod, dims, kind)
        R, _ = ref_matrix(ref, m, c)
        b = R.op("amul", 0, meshmod.cell_field_global(m, 42))
        levels, fwd = ref.reference_hierarchy(m.nCells, m.lower, m.upper, meshmod.face_area_pair_weights(m), 10,
                                              c["diag"], c["upper"], c["lower"])
        z = np.zeros(m.nCells)
        psi, p = ref.gamg_solve_levels(levels, z, b, **ctl)
        out[f"{name}.psi"] = psi
        out[f"{name}.perf"] = np.array([p["nIterations"], p["converged"], p["initialResidual"], p["finalResidual"]])
        hist = []
        for k in range(1, min(p["nIterations"], HIST_K) + 1):
            kw = dict(ctl)
            kw.update(tolerance=0.0, maxIter=k)
            hist.append(ref.gamg_solve_levels(levels, z, b, **kw)[1]["finalResidual"])
        out[f"{name}.hist"] = np.array(hist)
        out[f"{name}.levels"] = np.array([[lv["nCells"], len(lv["lower"])] for lv in levels[1:]])
        out[f"{name}.forward"] = np.array([fwd])
        for k in range(min(len(levels) - 1, 3)):
            out[f"{name}.restrict{k}"] = levels[k]["restrict"].astype(np.int32)
            out[f"{name}.coarseUpperAddr{k}"] = levels[k + 1]["upper"].astype(np.int32)
            out[f"{name}.coarseDiag{k}"] = levels[k + 1]["diag"]
    return out


if __name__ == "__main__":
    meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
    from oracle import ref_ldu
    assert ref_ldu.available(), "oracle/_ref is not built: this script needs /root/reference"
    data = generate(meshmod, ref_ldu)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path), "bytes,", len(data), "arrays")
