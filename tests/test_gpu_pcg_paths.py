"""PCG has two execution paths on the GPU: the reference's op list (7 launches, 3 global
sums per iteration; B200LDU_PCG_FUSED=0) and the fused form (4 launches, 2 sums; default).
Both must follow the oracle's residual history and agree with each other."""
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    capi = importlib.import_module("rapidcfd-dev_b200.capi")
    ctx = capi.Context(0)
    yield capi, ctx, torch
    ctx.close()


@pytest.mark.parametrize("pre", ["none", "diagonal", "DIC"])
@pytest.mark.parametrize("dims", [(18, 14, 10), (32, 32, 32)])
def test_fused_and_unfused_match_oracle(gpu, meshmod, orc, pre, dims):
    capi, ctx, torch = gpu
    m = meshmod.hex_mesh(*dims)
    c = meshmod.pressure_laplacian(m)
    oa = orc.Addr(m.nCells, m.lower, m.upper)
    om = orc.Matrix(oa, c["diag"], c["upper"], None)
    xs = meshmod.cell_field_global(m, 42)
    b = om.amul(xs)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    addr = capi.mesh_to_device(ctx, m)
    mat = capi.LduMatrix(addr)
    dg, up = t(c["diag"]), t(c["upper"])
    mat.set(dg, up)
    bd = t(b)
    res = {}
    for ctl in (dict(tolerance=1e-8, maxIter=400), dict(tolerance=0.0, maxIter=9),
                dict(tolerance=1e30, maxIter=50, minIter=4), dict(tolerance=1e30, maxIter=50)):
        psi_ref, pr, href = om.solve("PCG", pre, np.zeros(m.nCells), b, **ctl)
        for fused in ("0", "1"):
            os.environ["B200LDU_PCG_FUSED"] = fused
            psi = torch.zeros(m.nCells, dtype=torch.float64, device=ctx.device)
            perf, hist = mat.solve("PCG", pre, psi, bd, histCap=512, **ctl)
            assert abs(perf.nIterations - pr.nIterations) <= (0 if pr.nIterations < 100 else 1), (fused, ctl)
            assert perf.converged == pr.converged
            k = min(30, len(hist), len(href))
            np.testing.assert_allclose(hist[:k], href[:k], rtol=1e-9)
            assert len(hist) == perf.nIterations + 1
            np.testing.assert_allclose(psi.cpu().numpy(), psi_ref, rtol=0, atol=2e-7)
            res[fused] = (perf.nIterations, hist.copy())
        assert abs(res["0"][0] - res["1"][0]) <= (0 if pr.nIterations < 100 else 1)
    os.environ.pop("B200LDU_PCG_FUSED", None)
    mat.close()
    addr.close()
