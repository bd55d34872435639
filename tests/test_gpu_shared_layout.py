"""The optional shared-coefficient layout (B200LDU_SHARED=1; symmetric matrices, one stored
coefficient per face, cp.async double-buffered streams) must give the same bits as the
default per-entry layout for the row-gather kernels and the same PCG history."""
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dims,band", [((16, 16, 16), None), ((20, 13, 9), 128), ((32, 32, 32), None)])
def test_shared_layout_matches(meshmod, orc, dims, band):
    import torch
    capi = importlib.import_module("rapidcfd-dev_b200.capi")
    ctx = capi.Context(0)
    m = meshmod.hex_mesh(*dims)
    c = meshmod.pressure_laplacian(m)
    oa = orc.Addr(m.nCells, m.lower, m.upper)
    om = orc.Matrix(oa, c["diag"], c["upper"], None)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    x, b = meshmod.cell_field_global(m, 3), meshmod.cell_field_global(m, 4)
    xd, bd, dg, up = t(x), t(b), t(c["diag"]), t(c["upper"])
    if band:
        os.environ["B200LDU_BAND_ROWS"] = str(band)
    addr = capi.mesh_to_device(ctx, m)
    os.environ.pop("B200LDU_BAND_ROWS", None)
    os.environ["B200LDU_SHARED"] = "1"
    try:
        mat = capi.LduMatrix(addr)
        mat.set(dg, up)
        assert np.array_equal(mat.Amul(xd).cpu().numpy(), om.amul(x))
        assert np.array_equal(mat.Tmul(xd).cpu().numpy(), om.tmul(x))
        assert np.array_equal(mat.sumA(xd).cpu().numpy(), om.sumA())
        assert np.array_equal(mat.residual(xd, bd).cpu().numpy(), om.residual(x, b))
        assert np.array_equal(mat.H(xd).cpu().numpy(), om.H(x))
        assert np.array_equal(mat.H1(xd).cpu().numpy(), om.H1())
        assert np.array_equal(mat.smooth("Jacobi", xd, bd, 2).cpu().numpy(), om.jacobi(x, b, 2))
        ref = om.precondition("DIC", x)
        np.testing.assert_allclose(mat.precondition("DIC", xd).cpu().numpy(), ref, rtol=1e-13,
                                   atol=1e-13 * np.abs(ref).max())
        xs = meshmod.cell_field_global(m, 42)
        rhs = om.amul(xs)
        _, pr, href = om.solve("PCG", "DIC", np.zeros(m.nCells), rhs, tolerance=1e-8, maxIter=400)
        for fused in ("0", "1"):
            os.environ["B200LDU_PCG_FUSED"] = fused
            psi = torch.zeros(m.nCells, dtype=torch.float64, device=ctx.device)
            perf, hist = mat.solve("PCG", "DIC", psi, t(rhs), histCap=512, tolerance=1e-8, maxIter=400)
            assert abs(perf.nIterations - pr.nIterations) <= (0 if pr.nIterations < 100 else 1)
            k = min(30, len(hist))
            np.testing.assert_allclose(hist[:k], href[:k], rtol=1e-9)
            np.testing.assert_allclose(psi.cpu().numpy(), xs, rtol=0, atol=1e-5)
        mat.close()
    finally:
        os.environ.pop("B200LDU_SHARED", None)
        os.environ.pop("B200LDU_PCG_FUSED", None)
    addr.close()
    ctx.close()
