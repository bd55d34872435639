"""b200ldu_fv_limiter / b200ldu_fv_limited_weights on the device against the oracle (bit for bit), and an upwind
convection matrix assembled with them."""
import importlib

import numpy as np
import pytest

from oracle import limiters_oracle as lo
from test_limiters_cpu import _case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    capi = importlib.import_module("rapidcfd-dev_b200.capi")
    ctx = capi.Context(0)
    yield capi, ctx, torch
    ctx.close()


@pytest.mark.parametrize("scheme,k", [("upwind", 1.0), ("linear", 1.0), ("limitedLinear", 1.0), ("limitedLinear", 0.2),
                                      ("vanLeer", 1.0), ("Minmod", 1.0)])
def test_limiter_and_weights_on_the_device(gpu, meshmod, orc, scheme, k):
    capi, ctx, torch = gpu
    m, vf, gradc, flux, cd, cc = _case(meshmod, (12, 9, 7), 8)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    addr = capi.mesh_to_device(ctx, m)
    lim = capi.fv_limiter(addr, scheme, t(flux), t(vf), t(gradc), t(cc), k)
    want = lo.limiter(scheme, m.lower, m.upper, flux, vf, gradc, cc, k)
    assert np.array_equal(lim.cpu().numpy(), want, equal_nan=True)
    w = capi.fv_limited_weights(ctx, t(flux), lim, t(cd))
    assert np.array_equal(w.cpu().numpy(), lo.limited_weights(flux, want, cd), equal_nan=True)
    wu = capi.fv_limited_weights(ctx, t(flux))
    assert np.array_equal(wu.cpu().numpy(), lo.limited_weights(flux))
    # fvm::div(phi, .) with these weights: same coefficients as the oracle's fill
    lower, upper, diag = capi.fv_convection_fill(addr, w, t(flux))
    oa = orc.Addr(m.nCells, m.lower, m.upper)
    ol, ou, od = orc.convection_fill(oa, lo.limited_weights(flux, want, cd), flux)
    assert np.array_equal(lower.cpu().numpy(), ol) and np.array_equal(upper.cpu().numpy(), ou)
    assert np.array_equal(diag.cpu().numpy(), od)
    with pytest.raises(Exception, match="Unknown discretisation scheme"):
        capi.fv_limiter(addr, "QUICKEST", t(flux), t(vf), t(gradc), t(cc))
    addr.close()
