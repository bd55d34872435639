"""Generates tests/golden/ldu_golden.npz -- SELF-GENERATED fixtures.

The reference (RapidCFD-dev) ships no tests, fixtures or golden vectors and cannot be built in
this image, so these vectors do NOT pin the oracle against the reference:
they freeze the oracle's own outputs (commit that introduced this file) on seeded inputs, so that
 * `pytest -m "not gpu"` detects any silent drift of the oracle (bit-exact comparison), and
 * `pytest -m gpu` compares the CUDA path with vectors that exist independently of the oracle
   build on the GPU box.
Inputs are regenerated from the seeds by rapidcfd-dev_b200/mesh.py; only outputs are stored.

    python tests/golden/make_golden.py          # rewrites ldu_golden.npz
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

# (name, dims, kind) of the matrix-operation cases; same dims/kinds as tests/test_gpu_parity.py CASES
OPS_CASES = [("opsP", (13, 7, 5), "P"), ("opsU", (13, 7, 5), "U")]
# (name, dims, kind, solver, preconditioner/smoother, controls)
SOLVE_CASES = [
    ("pcgDIC", (16, 16, 16), "P", "PCG", "DIC", dict(tolerance=1e-7, maxIter=400)),
    ("pcgDiag", (16, 16, 16), "P", "PCG", "diagonal", dict(tolerance=1e-7, maxIter=400)),
    ("pcgNone", (20, 12, 9), "P", "PCG", "none", dict(tolerance=1e-7, maxIter=400)),
    ("pbicgDILU", (16, 12, 10), "U", "PBiCG", "DILU", dict(tolerance=1e-8, maxIter=300)),
    ("pbicgstabDILU", (16, 12, 10), "U", "PBiCGStab", "DILU", dict(tolerance=1e-8, maxIter=300)),
    ("smooth2", (10, 10, 10), "U", "smoothSolver", "GaussSeidel", dict(tolerance=1e-6, maxIter=500, nSweeps=2)),
]
GAMG_CASES = [("gamgP", (16, 14, 12), "P", dict(tolerance=1e-8, maxIter=100)),
              ("gamgU", (16, 14, 12), "U", dict(tolerance=1e-8, maxIter=100))]


def matrix_case(meshmod, orc, dims, kind):
    m = meshmod.hex_mesh(*dims)
    c = meshmod.pressure_laplacian(m) if kind == "P" else meshmod.momentum_matrix(m)
    a = orc.Addr(m.nCells, m.lower, m.upper)
    M = orc.Matrix(a, c["diag"], c["upper"], c["lower"])
    return m, c, a, M


def generate(meshmod, orc):
    out = {}
    for name, dims, kind in OPS_CASES:
        m, c, a, M = matrix_case(meshmod, orc, dims, kind)
        x, b = meshmod.cell_field_global(m, 3), meshmod.cell_field_global(m, 4)
        out[f"{name}.amul"] = M.amul(x)
        out[f"{name}.tmul"] = M.tmul(x)
        out[f"{name}.sumA"] = M.sumA()
        out[f"{name}.residual"] = M.residual(x, b)
        out[f"{name}.H"] = M.H(x)
        out[f"{name}.H1"] = M.H1()
        out[f"{name}.faceH"] = M.faceH(x)
        for pre in ("diagonal", "DIC"):
            for T in (False, True):
                out[f"{name}.pre.{pre}.{int(T)}"] = M.precondition(pre, x, T)
        for ns in (1, 3):
            out[f"{name}.jacobi{ns}"] = M.jacobi(x, b, ns)
    for name, dims, kind, solver, pre, ctl in SOLVE_CASES:
        m, c, a, M = matrix_case(meshmod, orc, dims, kind)
        xs = meshmod.cell_field_global(m, 42)
        b = M.amul(xs)
        psi, perf, hist = M.solve(solver, pre, np.zeros(m.nCells), b, **ctl)
        out[f"{name}.hist"] = np.asarray(hist)
        out[f"{name}.psi"] = psi
        out[f"{name}.perf"] = np.array([perf.nIterations, perf.converged, perf.initialResidual, perf.normFactor])
    for name, dims, kind, ctl in GAMG_CASES:
        m, c, a, M = matrix_case(meshmod, orc, dims, kind)
        g = orc.Gamg(a, meshmod.face_area_pair_weights(m), 10)
        xs = meshmod.cell_field_global(m, 42)
        b = M.amul(xs)
        psi, perf, hist = g.solve(M, "GaussSeidel", np.zeros(m.nCells), b, **ctl)
        out[f"{name}.hist"] = np.asarray(hist)
        out[f"{name}.psi"] = psi
        out[f"{name}.perf"] = np.array([perf.nIterations, perf.converged, perf.initialResidual, perf.normFactor])
        out[f"{name}.levels"] = np.array([[g.ncells(k), g.nfaces(k)] for k in range(g.nLevels)])
        for k in range(min(g.nLevels, 3)):
            out[f"{name}.restrict{k}"] = g.restrict_addr(k).astype(np.int32)
    return out


if __name__ == "__main__":
    meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
    from oracle import ldu_oracle as orc
    orc.build()
    data = generate(meshmod, orc)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ldu_golden.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path), "bytes,", len(data), "arrays")
