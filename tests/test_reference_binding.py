"""The reference-side binding of the product (rapidcfd-dev_b200/foam/b200Solver.H) COMPILED together with the
reference's own lduMatrixSolver.C / PCG.C / PBiCG.C ... (oracle/ref_harness/harness_binding.cpp ->
oracle/_ref/libref_binding.so): the reference's lduMatrix::solver::New reads `solver b200PCG;` from the
dictionary, finds the type in its run-time selection tables and dispatches into libb200ldu.so.
CPU: the table look-ups (no compute).  GPU: same solverPerformance as the reference's own solver."""
import importlib

import numpy as np
import pytest

from oracle import ref_ldu

pytestmark = pytest.mark.skipif(not ref_ldu.binding_available(), reason="oracle/_ref/libref_binding.so not built")


def _case(meshmod, dims, kind):
    m = meshmod.hex_mesh(*dims)
    c = meshmod.pressure_laplacian(m) if kind == "P" else meshmod.momentum_matrix(m)
    os_, ls, lo = ref_ldu.ldu_arrays(m.nCells, m.lower, m.upper)
    fixed = (m.nCells, m.lower, m.upper, os_, ls, lo, c["diag"], c["upper"], c["lower"])
    return m, c, fixed


def test_binding_is_selected_through_the_references_tables(meshmod):
    """no GPU here: construction through solver::New succeeds for the registered words (and only for the matrix
    kind they are registered for, like the reference's own: PCG.C:36, PBiCG.C:36), and solve() fails loudly"""
    m, c, fixed = _case(meshmod, (6, 5, 4), "P")
    z, b = np.zeros(m.nCells), np.ones(m.nCells)
    with pytest.raises(ValueError, match="unknown solver"):
        ref_ldu.solve("b200Nonsense", "DIC", *fixed, z, b, binding=True)
    with pytest.raises(ValueError, match="unknown solver"):     # asymmetric-only word on a symmetric matrix
        ref_ldu.solve("b200PBiCG", "DILU", *fixed, z, b, binding=True)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ref_ldu.solve("b200PCG", "DIC", *fixed, z, b, binding=True)
    # the reference's own solvers are in the same tables of the same library
    _, p = ref_ldu.solve("PCG", "DIC", *fixed, z, c["diag"] * 0 + 1.0, binding=True, maxIter=5)
    assert p["solverName"] == "AINVPCG"


@pytest.mark.gpu
@pytest.mark.parametrize("dims,kind,solver,pre", [((16, 12, 10), "P", "PCG", "DIC"), ((16, 12, 10), "P", "PCG", "diagonal"),
                                                  ((12, 10, 8), "U", "PBiCG", "DILU"), ((12, 10, 8), "U", "PBiCGStab", "DILU"),
                                                  ((10, 10, 10), "U", "smoothSolver", "GaussSeidel")])
def test_reference_dispatches_into_the_cuda_library(meshmod, orc, dims, kind, solver, pre):
    m, c, fixed = _case(meshmod, dims, kind)
    oa = orc.Addr(m.nCells, m.lower, m.upper)
    b = orc.Matrix(oa, c["diag"], c["upper"], c["lower"]).amul(meshmod.cell_field_global(m, 42))
    z = np.zeros(m.nCells)
    ctl = dict(tolerance=1e-9, maxIter=300, nSweeps=2)
    psi_ref, pr = ref_ldu.solve(solver, pre, *fixed, z, b, binding=True, **ctl)          # the reference's own class
    psi_gpu, pg = ref_ldu.solve("b200" + solver, pre, *fixed, z, b, binding=True, **ctl)  # b200Solver -> libb200ldu.so
    assert pg["solverName"] == pr["solverName"]
    assert pg["converged"] == pr["converged"] and pg["singular"] == pr["singular"]
    if solver == "PBiCGStab":   # the reference's yA/zA slip (PBiCGStab.C:263-270) is not reproduced by default
        assert pg["converged"]
        np.testing.assert_allclose(psi_gpu, meshmod.cell_field_global(m, 42), atol=1e-6)
        return
    assert abs(pg["nIterations"] - pr["nIterations"]) <= 1
    assert abs(pg["initialResidual"] - pr["initialResidual"]) <= 1e-12 * pr["initialResidual"]
    np.testing.assert_allclose(psi_gpu, psi_ref, rtol=0, atol=1e-7)
