"""The k-epsilon transport step ("two PBiCG scalar solves per step through the same boundary"):
the production-term kernel on the host against the oracle, the sequencing of rapidcfd-dev_b200/kepsilon.py over the oracle-backed
stand-in against oracle/kepsilon_oracle.py (bit for bit: same operations in the same order), and what the model is for."""
import ctypes as C
import importlib

import numpy as np
import pytest

import oracle_backend
from oracle import kepsilon_oracle as ko
from oracle import piso_oracle as po
from test_host_kernels_cpu import _d, hk  # noqa: F401  (fixture: the host build of the kernels)


def test_symm_magsqr_on_the_host(hk):  # noqa: F811
    rng = np.random.default_rng(5)
    T = np.ascontiguousarray(rng.uniform(-2, 2, (1000, 9)))
    T[:10] = 0.0
    out = np.zeros(1000)
    hk.hk_field_symm_magsqr.argtypes = [C.c_longlong, C.c_void_p, C.c_void_p]
    hk.hk_field_symm_magsqr(1000, _d(T), _d(out))
    assert np.array_equal(out, ko.symm_magsqr(T))
    S = 0.5 * (T.reshape(-1, 3, 3) + T.reshape(-1, 3, 3).transpose(0, 2, 1))
    assert np.allclose(out, (S * S).sum((1, 2)), rtol=1e-14)


def _both(meshmod, orc, n=6, steps=3):
    """the lid-driven cavity after a few PISO steps, on the stand-in and on the oracle, with k-epsilon objects on both"""
    ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
    kem = importlib.import_module("rapidcfd-dev_b200.kepsilon")
    capi, ctx, torch = oracle_backend.fixture()
    m, case = ico.cavity(capi, ctx, torch, n)
    _, ref = po.cavity_from_hex(orc, meshmod, n)
    ctl = dict(tolerance=1e-12, relTol=0.0)
    for _ in range(steps):
        case.step(UControls=ctl, pControls=ctl)
        ref.step(UControls=ctl, pControls=ctl)
    nB = len(ref.bfc)
    rng = np.random.default_rng(3)
    k0, e0 = rng.uniform(0.01, 0.02, m.nCells), rng.uniform(0.05, 0.1, m.nCells)
    k0[:5] = -1e-3                                     # bounded before the first use
    kB, eB = np.full(nB, 0.015), np.full(nB, 0.08)
    bMagSf, bDelta = np.full(nB, m.h * m.h), np.full(nB, 2.0 / m.h)
    dev = kem.KEpsilon(capi, case, bMagSf, bDelta, k0, e0, kB, eB)
    orf = ko.KEpsilon(orc, ref.addr, ref.Sf, ref.magSf, ref.w, ref.delta, ref.V, ref.bfc, ref.bSf, bMagSf, bDelta, ref.Ub, ref.nu, k0, e0,
                      kB, eB)
    return m, case, ref, dev, orf


@pytest.mark.parametrize("divScheme,alpha", [("upwind", None), ("linear", 0.7)])
def test_kepsilon_sequencing_over_the_oracle_backend(meshmod, orc, divScheme, alpha):
    m, case, ref, dev, orf = _both(meshmod, orc)
    assert np.array_equal(dev.k.numpy(), orf.k) and (orf.k > 0).all()             # bound() in the constructor
    assert np.array_equal(dev.nut.numpy(), orf.nut)
    for _ in range(2):
        pe, pk = dev.correct(case.U, case.phi, case.bphi, case.deltaT, divScheme, alpha, alpha)
        qe, qk = orf.correct(ref.U, ref.phi, ref.bphi, ref.deltaT, divScheme, alpha, alpha)
        assert pe.nIterations == qe.nIterations and pk.nIterations == qk.nIterations
        for a, b in ((dev.G, orf.G), (dev.epsilon, orf.epsilon), (dev.k, orf.k), (dev.nut, orf.nut)):
            assert np.array_equal(a.numpy(), b)
    assert (orf.k > 0).all() and (orf.epsilon > 0).all() and orf.G.max() > 0


def test_kepsilon_behaves_like_the_model(meshmod, orc):
    """no shear, no flux: k and epsilon decay as dk/dt = -epsilon, d(epsilon)/dt = -C2 epsilon^2/k (implicit Euler), uniformly"""
    m, case, ref, dev, orf = _both(meshmod, orc, steps=0)
    n = m.nCells
    orf.k, orf.epsilon = np.full(n, 0.02), np.full(n, 0.1)
    orf.kB, orf.epsB = np.full(len(ref.bfc), 0.02), np.full(len(ref.bfc), 0.1)
    orf.update_nut()
    dt = 0.01
    U0, phi0, b0 = np.zeros((n, 3)), np.zeros(len(ref.phi)), np.zeros(len(ref.bfc))
    orf.Ub = np.zeros_like(orf.Ub)
    orf.correct(U0, phi0, b0, dt, controls=dict(tolerance=1e-14, relTol=0.0))
    # interior cells away from the fixed-value walls follow the ODEs of homogeneous decaying turbulence
    inner = np.all((m.cell_centres() > 2.5 * m.h) & (m.cell_centres() < 1 - 2.5 * m.h), axis=1)
    e1 = 0.1 / (1 + dt * 1.92 * 0.1 / 0.02)             # (1/dt + C2 eps0/k0) eps1 = eps0/dt
    k1 = 0.02 / (1 + dt * e1 / 0.02)                    # (1/dt + eps1/k0) k1 = k0/dt
    assert np.allclose(orf.epsilon[inner], e1, rtol=1e-3) and np.allclose(orf.k[inner], k1, rtol=1e-3)
    assert orf.G.max() == 0.0
