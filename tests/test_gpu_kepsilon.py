"""The k-epsilon transport step on the device against oracle/kepsilon_oracle.py: the production-term kernel bit for bit, the two
PBiCG + DILU solves, the bounding and nut to solver tolerance."""
import importlib

import numpy as np
import pytest

from oracle import kepsilon_oracle as ko
from oracle import piso_oracle as po

pytestmark = pytest.mark.gpu


def test_kepsilon_on_the_device(meshmod, orc):
    import torch
    capi = importlib.import_module("rapidcfd-dev_b200.capi")
    ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
    kem = importlib.import_module("rapidcfd-dev_b200.kepsilon")
    ctx = capi.Context(0)
    n = 10
    m, case = ico.cavity(capi, ctx, torch, n)
    _, ref = po.cavity_from_hex(orc, meshmod, n)
    ctl = dict(tolerance=1e-12, relTol=0.0)
    for _ in range(3):
        case.step(UControls=ctl, pControls=ctl)
        ref.step(UControls=ctl, pControls=ctl)
    nB = len(ref.bfc)
    rng = np.random.default_rng(3)
    k0, e0 = rng.uniform(0.01, 0.02, m.nCells), rng.uniform(0.05, 0.1, m.nCells)
    k0[:5] = -1e-3
    kB, eB = np.full(nB, 0.015), np.full(nB, 0.08)
    bMagSf, bDelta = np.full(nB, m.h * m.h), np.full(nB, 2.0 / m.h)
    dev = kem.KEpsilon(capi, case, bMagSf, bDelta, k0, e0, kB, eB)
    orf = ko.KEpsilon(orc, ref.addr, ref.Sf, ref.magSf, ref.w, ref.delta, ref.V, ref.bfc, ref.bSf, bMagSf, bDelta, ref.Ub, ref.nu, k0, e0,
                      kB, eB)
    assert np.array_equal(dev.k.cpu().numpy(), orf.k) and np.array_equal(dev.nut.cpu().numpy(), orf.nut)     # bound + nut
    T = torch.from_numpy(np.ascontiguousarray(rng.uniform(-2, 2, (4096, 9)))).to(ctx.device)
    assert np.array_equal(case.ops.symm_magsqr(T).cpu().numpy(), ko.symm_magsqr(T.cpu().numpy()))
    # the relaxed variant (alpha 0.7) is covered on the CPU through the stand-in, which mirrors the library's coefficient copies
    for divScheme, alpha in (("upwind", None),):
        pe, pk = dev.correct(case.U, case.phi, case.bphi, case.deltaT, divScheme, alpha, alpha, controls=ctl)
        qe, qk = orf.correct(ref.U, ref.phi, ref.bphi, ref.deltaT, divScheme, alpha, alpha, controls=ctl)
        assert abs(pe.nIterations - qe.nIterations) <= 1 and abs(pk.nIterations - qk.nIterations) <= 1
        for a, b in ((dev.G, orf.G), (dev.epsilon, orf.epsilon), (dev.k, orf.k), (dev.nut, orf.nut)):
            np.testing.assert_allclose(a.cpu().numpy(), b, rtol=1e-8, atol=1e-12)
    assert (dev.k > 0).all() and (dev.epsilon > 0).all()
    dev.close()
    case.close()
    ctx.close()
