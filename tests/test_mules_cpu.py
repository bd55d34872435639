"""Explicit MULES: the oracle against the reference's own MULESTemplates.C compiled here
(oracle/_ref/libref_mules.so: limiter, limit, explicitSolve with the reference's one / zero field algebra), the device code
of csrc/mules_kernels.cuh executed on the host against the oracle, and the sequencing of rapidcfd-dev_b200/mules.py over
the oracle-backed stand-in -- all bit for bit; then what the limiter is for: a bounded, conservative update."""
import ctypes as C
import importlib

import numpy as np
import pytest

from oracle import mules_oracle as mo
from test_host_kernels_cpu import Host, _d, hk  # noqa: F401  (fixture: the host build of the kernels)

COMBOS = ["one-zero", "rho", "SpSu", "rho-SpSu"]


def case(meshmod, dims=(7, 5, 4), seed=1, combo="one-zero", scale=1e-3):
    m = meshmod.hex_mesh(*dims)
    ps, bfc = m.patch_start_facecells(m.wall_patches())
    rng = np.random.default_rng(seed)
    n, nF, nB = m.nCells, m.nFaces, len(bfc)
    d = dict(m=m, ps=ps, bfc=bfc, n=n, nF=nF, nB=nB, rDeltaT=50.0)
    d["V"] = m.volumes() * rng.uniform(0.8, 1.2, n)
    d["psi"] = rng.uniform(0, 1, n)
    d["psi"][: n // 5] = 1.0                       # saturated cells: no room upwards
    d["psi"][n // 5: n // 3] = 0.0
    d["psi0"] = d["psi"].copy()
    d["psiB"] = rng.uniform(0, 1, nB)
    d["phi"], d["phiB"] = rng.uniform(-1, 1, nF) * scale, rng.uniform(-1, 1, nB) * scale
    d["phiB"][: nB // 3] = 0.0                     # walls
    d["phi"][::13] = 0.0                           # pos(0) = 1; phiCorr = 0 goes down the `else` branches
    d["phiPsi"] = d["phi"] * (0.5 * (d["psi"][m.lower] + d["psi"][m.upper]))        # central: unbounded
    d["phiPsiB"] = d["phiB"] * d["psiB"]
    kw = {}
    if "rho" in combo:
        kw.update(rho=rng.uniform(0.9, 1.1, n), rho0=rng.uniform(0.9, 1.1, n))
    if "SpSu" in combo:
        kw.update(Sp=-rng.uniform(0, 1, n), Su=rng.uniform(0, 0.1, n))
    d["kw"] = kw
    return d


def _cat(a, b):
    return np.concatenate([a, b])


@pytest.mark.skipif(not mo.reference_available(), reason="oracle/_ref/libref_mules.so not built")
@pytest.mark.parametrize("combo", COMBOS)
@pytest.mark.parametrize("nIter", [0, 1, 3])
def test_oracle_matches_the_reference_mules(meshmod, combo, nIter):
    d = case(meshmod, combo=combo, seed=2 + nIter)
    m, kw = d["m"], d["kw"]
    bd, bdB = mo.upwind_flux(m.lower, m.upper, d["phi"], d["phiB"], d["psi"], d["psiB"])
    corr, corrB = d["phiPsi"] - bd, d["phiPsiB"] - bdB
    lam, lamB = mo.limiter(d["n"], m.lower, m.upper, d["bfc"], d["V"], d["rDeltaT"], d["psi"], d["psi0"], d["psiB"], bd, bdB, corr,
                           corrB, 1.0, 0.0, nIter, **kw)
    ref = mo.reference(0, d["n"], m.lower, m.upper, d["ps"], d["bfc"], d["V"], d["rDeltaT"], d["psi"], d["psi0"], d["psiB"],
                       _cat(bd, bdB), _cat(corr, corrB), 1.0, 0.0, nIter, **kw)
    assert np.array_equal(_cat(lam, lamB), ref)
    assert nIter == 0 or (lam < 1).sum() > d["nF"] // 4          # the limiter is active on this case
    lp, lpB = mo.limit(d["n"], m.lower, m.upper, d["bfc"], d["V"], d["rDeltaT"], d["psi"], d["psi0"], d["psiB"], d["phi"], d["phiB"],
                       d["phiPsi"], d["phiPsiB"], 1.0, 0.0, nIter, **kw)
    ref = mo.reference(1, d["n"], m.lower, m.upper, d["ps"], d["bfc"], d["V"], d["rDeltaT"], d["psi"], d["psi0"], d["psiB"],
                       _cat(d["phi"], d["phiB"]), _cat(d["phiPsi"], d["phiPsiB"]), 1.0, 0.0, nIter, **kw)
    assert np.array_equal(_cat(lp, lpB), ref)
    new = mo.explicit_solve(d["n"], m.lower, m.upper, d["bfc"], d["V"], d["rDeltaT"], d["psi0"], lp, lpB, **kw)
    ref = mo.reference(2, d["n"], m.lower, m.upper, d["ps"], d["bfc"], d["V"], d["rDeltaT"], d["psi"], d["psi0"], d["psiB"],
                       _cat(lp, lpB), None, **kw)
    assert np.array_equal(new, ref)


@pytest.mark.skipif(not mo.reference_available(), reason="oracle/_ref/libref_mules.so not built")
def test_oracle_matches_the_reference_other_bounds(meshmod):
    """psiMax / psiMin that cut into the field (the local extrema are clipped to them)"""
    d = case(meshmod, (5, 6, 3), seed=9)
    m = d["m"]
    bd, bdB = mo.upwind_flux(m.lower, m.upper, d["phi"], d["phiB"], d["psi"], d["psiB"])
    corr, corrB = d["phiPsi"] - bd, d["phiPsiB"] - bdB
    lam, lamB = mo.limiter(d["n"], m.lower, m.upper, d["bfc"], d["V"], d["rDeltaT"], d["psi"], d["psi0"], d["psiB"], bd, bdB, corr,
                           corrB, 0.8, 0.3, 2)
    ref = mo.reference(0, d["n"], m.lower, m.upper, d["ps"], d["bfc"], d["V"], d["rDeltaT"], d["psi"], d["psi0"], d["psiB"],
                       _cat(bd, bdB), _cat(corr, corrB), 0.8, 0.3, 2)
    assert np.array_equal(_cat(lam, lamB), ref)


@pytest.mark.parametrize("combo", COMBOS)
def test_device_mules_code_on_the_host(hk, meshmod, orc, combo):  # noqa: F811
    d = case(meshmod, (6, 7, 5), seed=4, combo=combo)
    m, kw, n, nF, nB = d["m"], d["kw"], d["n"], d["nF"], d["nB"]
    a = orc.Addr(n, m.lower, m.upper)
    H = Host(a, dict(bfc=d["bfc"], diag=np.zeros(n), upper=np.zeros(nF), lower=None))
    bd, bdB = mo.upwind_flux(m.lower, m.upper, d["phi"], d["phiB"], d["psi"], d["psiB"])
    corr, corrB = d["phiPsi"] - bd, d["phiPsiB"] - bdB
    f = lambda x: None if x is None else np.ascontiguousarray(x, np.float64)
    arrs = {k: f(v) for k, v in dict(psi=d["psi"], psi0=d["psi0"], psiB=d["psiB"], bd=bd, bdB=bdB, corr=corr, corrB=corrB, V=d["V"],
                                     rho=kw.get("rho"), rho0=kw.get("rho0"), Sp=kw.get("Sp"), Su=kw.get("Su")).items()}
    for nIter in (0, 2, 3):
        lam, lamB, scratch = np.ones(nF), np.ones(nB), np.zeros(6 * n)
        hk.hk_mules_limiter.argtypes = [C.c_void_p, C.c_int, C.c_double] + [C.c_void_p] * 12 + [C.c_double, C.c_double] + [C.c_void_p] * 3
        hk.hk_mules_limiter(H.p(), nIter, d["rDeltaT"], _d(arrs["rho"]), _d(arrs["rho0"]), _d(arrs["psi"]), _d(arrs["psi0"]),
                            _d(arrs["psiB"]), _d(arrs["bd"]), _d(arrs["bdB"]), _d(arrs["corr"]), _d(arrs["corrB"]), _d(arrs["Sp"]),
                            _d(arrs["Su"]), _d(arrs["V"]), 1.0, 0.0, _d(lam), _d(lamB), _d(scratch))
        want, wantB = mo.limiter(n, m.lower, m.upper, d["bfc"], d["V"], d["rDeltaT"], d["psi"], d["psi0"], d["psiB"], bd, bdB, corr,
                                 corrB, 1.0, 0.0, nIter, **kw)
        assert np.array_equal(lam, want) and np.array_equal(lamB, wantB)


@pytest.mark.parametrize("combo", COMBOS)
def test_mules_sequencing_over_the_oracle_backend(meshmod, orc, combo):
    """rapidcfd-dev_b200/mules.py issues the reference's operations in the reference's order"""
    import torch
    import oracle_backend as ob
    mules = importlib.import_module("rapidcfd-dev_b200.mules")
    capi, ctx, _ = ob.fixture()
    d = case(meshmod, (5, 4, 6), seed=6, combo=combo)
    m, kw = d["m"], d["kw"]
    addr = capi.mesh_to_device(ctx, m)
    capi.fv_boundary_set(addr, d["bfc"])
    ops = capi.FieldOps(ctx)
    t = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x, np.float64))
    tk = {k: t(v) for k, v in kw.items()}
    lp, lpB = mules.limit(capi, addr, ops, t(d["V"]), d["rDeltaT"], t(d["psi"]), t(d["psi0"]), t(d["psiB"]), t(d["phi"]), t(d["phiB"]),
                          t(d["phiPsi"]), t(d["phiPsiB"]), 1.0, 0.0, 3, **tk)
    want, wantB = mo.limit(d["n"], m.lower, m.upper, d["bfc"], d["V"], d["rDeltaT"], d["psi"], d["psi0"], d["psiB"], d["phi"], d["phiB"],
                           d["phiPsi"], d["phiPsiB"], 1.0, 0.0, 3, **kw)
    assert np.array_equal(lp.numpy(), want) and np.array_equal(lpB.numpy(), wantB)
    new = mules.explicit_solve(capi, addr, ops, t(d["V"]), d["rDeltaT"], t(d["psi0"]), lp, lpB, **tk)
    assert np.array_equal(new.numpy(), mo.explicit_solve(d["n"], m.lower, m.upper, d["bfc"], d["V"], d["rDeltaT"], d["psi0"], want,
                                                         wantB, **kw))


def advect(meshmod, limited, steps=12):
    """a slab of psi = 1 carried along x through a closed box by a solenoidal (uniform, walls closed) flux, central face values"""
    m = meshmod.hex_mesh(16, 3, 3)
    ps, bfc = m.patch_start_facecells(m.wall_patches())
    n, nB = m.nCells, len(bfc)
    V = m.volumes()
    x = m.cell_centres()[:, 0]
    psi = np.where((x > 0.2) & (x < 0.5), 1.0, 0.0)
    u, dt = 1.0, 0.3 * m.h                                     # Courant 0.3
    phi = np.where(m.faceDir == 0, u * m.h * m.h, 0.0)         # uniform in x ...
    interior = (m.lower % m.nx) < m.nx - 2                     # ... stopped two cells before the end wall:
    phi = np.where(interior, phi, 0.0)                         # the slab piles up against a flux-free face
    phiB, psiB = np.zeros(nB), np.zeros(nB)
    lo, hi, mass = [], [], []
    for _ in range(steps):
        phiPsi, phiPsiB = phi * (0.5 * (psi[m.lower] + psi[m.upper])), phiB * psiB
        if limited:
            phiPsi, phiPsiB = mo.limit(n, m.lower, m.upper, bfc, V, 1 / dt, psi, psi, psiB, phi, phiB, phiPsi, phiPsiB, 1.0, 0.0, 3)
        psi = mo.explicit_solve(n, m.lower, m.upper, bfc, V, 1 / dt, psi, phiPsi, phiPsiB)
        lo.append(psi.min()), hi.append(psi.max()), mass.append((psi * V).sum())
    return np.array(lo), np.array(hi), np.array(mass)


def test_limited_update_is_bounded_and_conservative(meshmod):
    lo, hi, mass = advect(meshmod, False)
    assert lo.min() < -0.05 and hi.max() > 1.05                # central differencing alone over- and undershoots
    lo, hi, massL = advect(meshmod, True)
    assert lo.min() >= -1e-12 and hi.max() <= 1 + 1e-12        # MULES keeps psi within [psiMin, psiMax]
    assert np.allclose(massL, massL[0], rtol=1e-13) and np.allclose(mass, massL[0], rtol=1e-13)   # fluxes only: conservative
