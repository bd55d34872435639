"""b200ldu_mules_limiter and the compositions of rapidcfd-dev_b200/mules.py on the device against the oracle (bit for bit),
and a 64^3 advection step through the same calls: bounded and conservative."""
import importlib

import numpy as np
import pytest

from oracle import mules_oracle as mo
from test_mules_cpu import COMBOS, case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    capi = importlib.import_module("rapidcfd-dev_b200.capi")
    ctx = capi.Context(0)
    yield capi, ctx, torch
    ctx.close()


@pytest.mark.parametrize("combo", COMBOS)
def test_mules_on_the_device(gpu, meshmod, combo):
    capi, ctx, torch = gpu
    mules = importlib.import_module("rapidcfd-dev_b200.mules")
    d = case(meshmod, (11, 9, 8), seed=12, combo=combo)
    m, kw = d["m"], d["kw"]
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(ctx.device)
    addr = capi.mesh_to_device(ctx, m)
    capi.fv_boundary_set(addr, d["bfc"])
    ops = capi.FieldOps(ctx)
    tk = {k: t(v) for k, v in kw.items()}
    bd, bdB = mo.upwind_flux(m.lower, m.upper, d["phi"], d["phiB"], d["psi"], d["psiB"])
    corr, corrB = d["phiPsi"] - bd, d["phiPsiB"] - bdB
    for nIter in (0, 1, 3):
        lam, lamB = capi.mules_limiter(addr, t(d["V"]), d["rDeltaT"], t(d["psi"]), t(d["psi0"]), t(d["psiB"]), t(bd), t(bdB), t(corr),
                                       t(corrB), 1.0, 0.0, nIter, tk.get("rho"), tk.get("rho0"), tk.get("Sp"), tk.get("Su"))
        want, wantB = mo.limiter(d["n"], m.lower, m.upper, d["bfc"], d["V"], d["rDeltaT"], d["psi"], d["psi0"], d["psiB"], bd, bdB, corr,
                                 corrB, 1.0, 0.0, nIter, **kw)
        assert np.array_equal(lam.cpu().numpy(), want) and np.array_equal(lamB.cpu().numpy(), wantB)
    lp, lpB = mules.limit(capi, addr, ops, t(d["V"]), d["rDeltaT"], t(d["psi"]), t(d["psi0"]), t(d["psiB"]), t(d["phi"]), t(d["phiB"]),
                          t(d["phiPsi"]), t(d["phiPsiB"]), 1.0, 0.0, 3, **tk)
    want, wantB = mo.limit(d["n"], m.lower, m.upper, d["bfc"], d["V"], d["rDeltaT"], d["psi"], d["psi0"], d["psiB"], d["phi"], d["phiB"],
                           d["phiPsi"], d["phiPsiB"], 1.0, 0.0, 3, **kw)
    assert np.array_equal(lp.cpu().numpy(), want) and np.array_equal(lpB.cpu().numpy(), wantB)
    new = mules.explicit_solve(capi, addr, ops, t(d["V"]), d["rDeltaT"], t(d["psi0"]), lp, lpB, **tk)
    assert np.array_equal(new.cpu().numpy(), mo.explicit_solve(d["n"], m.lower, m.upper, d["bfc"], d["V"], d["rDeltaT"], d["psi0"],
                                                               want, wantB, **kw))
    addr.close()


@pytest.mark.parametrize("combo", COMBOS)
def test_cmules_on_the_device(gpu, meshmod, combo):
    """b200ldu_mules_limiter_corr, limit_corr and correct against the oracle, bit for bit"""
    from test_mules_cpu import corr_case
    capi, ctx, torch = gpu
    mules = importlib.import_module("rapidcfd-dev_b200.mules")
    d = corr_case(meshmod, (11, 9, 8), 13, combo)
    m, kw = d["m"], d["kw"]
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(ctx.device)
    addr = capi.mesh_to_device(ctx, m)
    capi.fv_boundary_set(addr, d["bfc"])
    ops = capi.FieldOps(ctx)
    tk = {k: t(v) for k, v in kw.items()}
    for ex in (0.0, 0.2):
        lam, lamB = capi.mules_limiter_corr(addr, t(d["V"]), d["rDeltaT"], t(d["psi"]), t(d["psiB"]), t(d["phiB"]), t(d["corr"]), t(d["corrB"]),
                                            1.0, 0.0, 3, tk.get("rho"), tk.get("Sp"), tk.get("Su"), ex)
        want, wantB = mo.limiter(d["n"], m.lower, m.upper, d["bfc"], d["V"], d["rDeltaT"], d["psi"], d["psi"], d["psiB"], np.zeros(d["nF"]),
                                 d["phiB"], d["corr"], d["corrB"], 1.0, 0.0, 3, kw.get("rho"), None, kw.get("Sp"), kw.get("Su"), corr=True,
                                 extremaCoeff=ex)
        assert np.array_equal(lam.cpu().numpy(), want) and np.array_equal(lamB.cpu().numpy(), wantB)
    lc, lcB = mules.limit_corr(capi, addr, ops, t(d["V"]), d["rDeltaT"], t(d["psi"]), t(d["psiB"]), t(d["phiB"]), t(d["corr"]), t(d["corrB"]),
                               1.0, 0.0, 3, **tk)
    want, wantB = mo.limit_corr(d["n"], m.lower, m.upper, d["bfc"], d["V"], d["rDeltaT"], d["psi"], d["psiB"], d["phiB"], d["corr"],
                                d["corrB"], 1.0, 0.0, 3, **kw)
    assert np.array_equal(lc.cpu().numpy(), want) and np.array_equal(lcB.cpu().numpy(), wantB)
    new = mules.correct(capi, addr, ops, t(d["V"]), d["rDeltaT"], t(d["psi"]), lc, lcB, **tk)
    assert np.array_equal(new.cpu().numpy(), mo.correct(d["n"], m.lower, m.upper, d["bfc"], d["V"], d["rDeltaT"], d["psi"], want, wantB, **kw))
    addr.close()


def test_mules_advection_64_cubed_is_bounded_and_conservative(gpu, meshmod):
    """a disc of psi = 1 carried round by a discretely solenoidal flux (differences of a stream function that vanishes on the
    walls): central face values limited by MULES stay within [0, 1] to rounding and the total is conserved; without the limiter
    they leave the bounds by O(1)"""
    capi, ctx, torch = gpu
    mules = importlib.import_module("rapidcfd-dev_b200.mules")
    N = 64
    m = meshmod.hex_mesh(N)
    ps, bfc = m.patch_start_facecells(m.wall_patches())
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(ctx.device)
    addr = capi.mesh_to_device(ctx, m)
    capi.fv_boundary_set(addr, bfc)
    ops = capi.FieldOps(ctx)
    cc, h = m.cell_centres(), m.h
    psi0 = np.where(((cc[:, :2] - np.array([0.5, 0.3])) ** 2).sum(1) < 0.15 ** 2, 1.0, 0.0)
    S = lambda x, y: (np.sin(np.pi * x) * np.sin(np.pi * y)) ** 2 / np.pi
    i, j = m.lower % N, (m.lower // N) % N
    phix = (S((i + 1) * h, (j + 1) * h) - S((i + 1) * h, j * h)) * h
    phiy = -(S((i + 1) * h, (j + 1) * h) - S(i * h, (j + 1) * h)) * h
    phi = np.where(m.faceDir == 0, phix, np.where(m.faceDir == 1, phiy, 0.0))
    dt = 0.25 * h / (np.abs(phi).max() / h / h)                 # Courant 0.25
    V, phiD = t(m.volumes()), t(phi)
    zB = torch.zeros(len(bfc), dtype=torch.float64, device=ctx.device)
    w = t(m.weights())
    out = {}
    for limited in (False, True):
        psi = t(psi0)
        for _ in range(40):
            phiPsi = ops.mul(phiD, capi.fv_interpolate_linear(addr, 1, w, psi))
            phiPsiB = zB
            if limited:
                phiPsi, phiPsiB = mules.limit(capi, addr, ops, V, 1 / dt, psi, psi, zB, phiD, zB, phiPsi, zB, 1.0, 0.0, 3)
            psi = mules.explicit_solve(capi, addr, ops, V, 1 / dt, psi, phiPsi, phiPsiB)
        out[limited] = psi.cpu().numpy()
    assert out[False].min() < -0.05 and out[False].max() > 1.05
    assert out[True].min() >= -1e-10 and out[True].max() <= 1 + 1e-10
    for v in out.values():
        assert abs(v.sum() - psi0.sum()) <= 1e-9 * psi0.sum()
    addr.close()


@pytest.mark.parametrize("world", [2, 8])
def test_multi_gpu_mules(world):
    """MULES over processor patches, one rank per GPU (needs >= 2 GPUs; skipped otherwise)"""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29900 + world),
           os.path.join(root, "tests", "multi_gpu_mules_worker.py")]
    p = subprocess.run(cmd, cwd=root, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-4000:]
    assert p.stdout.count("MULTI-GPU-MULES-OK") == 4 * world
