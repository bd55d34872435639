"""Row a17 on the GPU: the fvMatrix glue (csrc/fvmatrix.cu through the C ABI) against oracle/fvm_oracle.py.
Everything that is a sum of separately rounded terms in the reference's order is compared bit for bit; the
solves to the tolerance of the Krylov parity tests.  Sorted last on purpose: these entry points were written
after the round's GPU budget was spent."""
import importlib
import os

import numpy as np
import pytest

from oracle import fvm_oracle as fo
from test_oracle_fvm import make, momentum_case, poisson_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    if os.environ.get("B200LDU_DRYRUN_ORACLE") == "1":   # CPU dry run of the tests' own logic: tests/oracle_backend.py
        import oracle_backend
        yield oracle_backend.fixture()
        return
    import torch
    assert torch.cuda.is_available()
    capi = importlib.import_module("rapidcfd-dev_b200.capi")
    ctx = capi.Context(0)
    yield capi, ctx, torch
    ctx.close()


class Dev:
    def __init__(self, gpu, m, d, nc, psi, addr=None, cou=None):
        capi, ctx, torch = gpu
        self.capi, self.torch, self.dev = capi, torch, ctx.device
        self.addr = addr if addr is not None else capi.mesh_to_device(ctx, m)
        bfc = np.ascontiguousarray(d["bfc"], dtype=np.int32)
        capi.fv_boundary_set(self.addr, bfc)
        self.t = lambda a: torch.from_numpy(np.array(a, dtype=np.float64).ravel()).to(ctx.device)   # always a copy
        self.diag, self.upper = self.t(d["diag"]), self.t(d["upper"])
        self.lower = self.t(d["lower"]) if d["lower"] is not None else None
        self.cou = None if cou is None else (self.t(cou[0]), self.t(cou[1]))
        self.mat = capi.LduMatrix(self.addr)
        self.set()
        self.source, self.psi = self.t(d["source"]), self.t(psi)
        self.fv = capi.FvMatrix(self.mat, nc, self.diag, self.source, self.psi, self.t(d["V"]), self.t(d["ic"]),
                                self.t(d["bc"]))
        self.nB = len(bfc)

    def set(self):
        if self.cou is None:
            self.mat.set(self.diag, self.upper, self.lower)
        else:
            self.mat.set(self.diag, self.upper, self.lower, self.cou[0], self.cou[1])   # bou, int

    def close(self):
        self.mat.close()
        self.addr.close()


def host(x, nc=None):
    a = x.cpu().numpy()
    return a if nc is None else a.reshape(-1, nc)


@pytest.mark.parametrize("nc", [1, 3])
def test_fvm_operations_bit_exact(gpu, meshmod, orc, nc):
    m, a, d = (poisson_case if nc == 1 else momentum_case)(meshmod, orc, (9, 7, 6))
    x = np.random.default_rng(2).uniform(-1, 1, (m.nCells, nc))
    ofm = make(orc, a, d, nc, x)
    D = Dev(gpu, m, d, nc, x)
    fv = D.fv
    assert np.array_equal(host(fv.A()), ofm.A())
    assert np.array_equal(host(fv.H(), nc), ofm.H())
    fi, fb, fc = fv.flux(D.nB)
    oi, ob, oc = ofm.flux()
    assert np.array_equal(host(fi, nc), oi) and np.array_equal(host(fb, nc), ob) and fc.numel() == 0
    for cmpt in list(range(nc)) + [-1]:
        got = D.t(d["diag"])
        if cmpt >= 0:
            fv.addBoundaryDiag(got, cmpt)
        else:
            fv.addCmptAvBoundaryDiag(got)
        ref = d["diag"].copy()
        ofm.addBoundaryDiag(ref, cmpt) if cmpt >= 0 else ofm.addCmptAvBoundaryDiag(ref)
        assert np.array_equal(host(got), ref), cmpt
    src = D.t(d["source"])
    fv.addBoundarySource(src)
    ref = np.array(d["source"], float).reshape(m.nCells, nc).copy()
    ofm.addBoundarySource(ref, False)
    assert np.array_equal(host(src, nc), ref)
    if nc == 1:
        assert np.array_equal(host(fv.residual()), ofm.residual())
    # relax and setReference work in place on diag / source
    for alpha in (1.0, 0.7):
        Dr = Dev(gpu, m, d, nc, x, addr=D.addr)
        orl = make(orc, a, d, nc, x)
        Dr.fv.relax(alpha)
        orl.relax(alpha)
        assert np.array_equal(host(Dr.diag), orl.diag) and np.array_equal(host(Dr.source, nc), orl.source)
        Dr.mat.close()
    val = [0.5, -1.25, 2.0][:nc]
    fv.setReference(11, val)
    fv.setReference(-1, val)
    ofm.setReference(11, val)
    assert np.array_equal(host(D.diag), ofm.diag) and np.array_equal(host(D.source, nc), ofm.source)
    D.close()


@pytest.mark.parametrize("nc,solver,pre", [(1, "PCG", "DIC"), (3, "PBiCG", "DILU"), (3, "smoothSolver", "GaussSeidel")])
def test_fvm_solve_segregated(gpu, meshmod, orc, nc, solver, pre):
    m, a, d = (poisson_case if nc == 1 else momentum_case)(meshmod, orc, (10, 8, 6))
    ctl = dict(tolerance=1e-10, maxIter=400)
    psi_ref, perfs_ref, _ = make(orc, a, d, nc).solve(solver, pre, **ctl)
    D = Dev(gpu, m, d, nc, np.zeros((m.nCells, nc)))
    perfs = D.fv.solve(solver, pre, **ctl)
    assert len(perfs) == nc
    for k in range(nc):
        assert perfs[k].converged == perfs_ref[k].converged
        assert abs(perfs[k].nIterations - perfs_ref[k].nIterations) <= 2
        assert perfs[k].solverName == perfs_ref[k].solverName
        np.testing.assert_allclose(perfs[k].initialResidual, perfs_ref[k].initialResidual, rtol=1e-10)
    np.testing.assert_allclose(host(D.psi, nc), psi_ref, rtol=0, atol=1e-8)
    # the matrix points at the caller's diagonal again (saveDiag): Amul with it equals the unfolded oracle matrix
    xs = meshmod.cell_field_global(m, 5)
    om = orc.Matrix(a, d["diag"], d["upper"], d["lower"])
    assert np.array_equal(host(D.mat.Amul(D.t(xs))), om.amul(xs))
    D.close()


def test_fvm_with_coupled_patches(gpu, meshmod, orc):
    """Cyclic patch pair on one device: the coupled branches of addBoundaryDiag / addBoundarySource / H / flux /
    residual / relax / solve, with the patchNeighbourField handed in as the reference's boundary condition would."""
    from test_oracle_core import _cyclic_case
    capi, ctx, torch = gpu
    m, c, ps, fc, nr, lo, hi = _cyclic_case(meshmod, "U")
    nc = 3
    rng = np.random.default_rng(8)
    wall = np.concatenate([p.faceCells for p in m.wall_patches() if p.name not in ("xmin", "xmax")]).astype(np.int32)
    value = rng.uniform(-1, 1, (len(wall), nc))
    ic, bc = fo.fixedValue_laplacian_coeffs(np.full(len(wall), 0.01 * m.h * m.h), np.full(len(wall), 2.0 / m.h), value)
    # the coupled internal coefficients enter through addBoundaryDiag: take them out of the assembled diagonal
    diag = c["diag"].copy()
    np.subtract.at(diag, fc, c["int"])
    d = dict(diag=diag, upper=c["upper"], lower=c["lower"], source=rng.uniform(-1, 1, (m.nCells, nc)) * m.h ** 3,
             bfc=wall, ic=-ic, bc=-bc, V=m.volumes())
    a = orc.Addr(m.nCells, m.lower, m.upper, ps, fc, neighbRank=nr)
    x = rng.uniform(-1, 1, (m.nCells, nc))
    kw = dict(couInt=c["int"], couBou=c["bou"])
    ofm = make(orc, a, d, nc, x, **kw)
    pnf = ofm.patchNeighbourField()
    assert np.array_equal(pnf, x[np.concatenate([hi, lo])])       # cyclic: the partner patch's cells
    addr = capi.LduAddressing(ctx, m.nCells, m.lower, m.upper, ps, fc, nr, m.cell_centres())
    D = Dev(gpu, m, d, nc, x, addr=addr, cou=(c["bou"], c["int"]))
    pnf_d = D.t(pnf)
    assert np.array_equal(host(D.fv.A()), ofm.A())
    assert np.array_equal(host(D.fv.H(pnf_d), nc), ofm.H())
    fi, fb, fcp = D.fv.flux(D.nB, len(fc), pnf_d)
    oi, ob, oc = ofm.flux()
    assert np.array_equal(host(fi, nc), oi) and np.array_equal(host(fb, nc), ob) and np.array_equal(host(fcp, nc), oc)
    src = D.t(d["source"])
    D.fv.addBoundarySource(src, pnf_d)
    ref = d["source"].copy()
    ofm.addBoundarySource(ref, True)
    assert np.array_equal(host(src, nc), ref)
    orl = make(orc, a, d, nc, x, **kw)
    Dr = Dev(gpu, m, d, nc, x, addr=addr, cou=(c["bou"], c["int"]))
    Dr.fv.relax(0.8)
    orl.relax(0.8)
    assert np.array_equal(host(Dr.diag), orl.diag) and np.array_equal(host(Dr.source, nc), orl.source)
    Dr.mat.close()
    # scalar residual with coupled patches (the doubled neighbour term included)
    d1 = dict(d, source=d["source"][:, 0].copy(), ic=d["ic"][:, :1].copy(), bc=d["bc"][:, :1].copy())
    o1 = make(orc, a, d1, 1, x[:, :1], **kw)
    D1 = Dev(gpu, m, d1, 1, x[:, :1], addr=addr, cou=(c["bou"], c["int"]))
    np.testing.assert_allclose(host(D1.fv.residual(D1.t(pnf[:, :1]))), o1.residual(), rtol=1e-13, atol=1e-14)
    D1.mat.close()
    # component loop: the coupled source goes in for all components and out again per component
    ctl = dict(tolerance=1e-10, maxIter=400)
    z = np.zeros((m.nCells, nc))
    oz = make(orc, a, d, nc, z, **kw)
    psi_ref, perfs_ref, _ = oz.solve("PBiCG", "DILU", **ctl)
    Dz = Dev(gpu, m, d, nc, z, addr=addr, cou=(c["bou"], c["int"]))
    perfs = Dz.fv.solve("PBiCG", "DILU", pnf=Dz.t(oz.patchNeighbourField()), **ctl)
    for k in range(nc):
        assert perfs[k].converged and abs(perfs[k].nIterations - perfs_ref[k].nIterations) <= 2
    np.testing.assert_allclose(host(Dz.psi, nc), psi_ref, rtol=0, atol=1e-8)
    Dz.mat.close()
    D.close()


def test_icofoam_step_matches_oracle(gpu, meshmod, orc):
    """the device icoFoam step (rapidcfd-dev_b200/icofoam.py over the C ABI) against the
    oracle's step (oracle/piso_oracle.py) on the lid-driven cavity.  Same operator sequence, so everything except the
    linear solves is bit for bit; the solves run to 1e-12 and the fields are compared at 1e-9."""
    from oracle import piso_oracle as po
    capi, ctx, torch = gpu
    ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
    n = 8
    m, ref = po.cavity_from_hex(orc, meshmod, n, nu=0.01)
    m2, dev = ico.cavity(capi, ctx, torch, n, nu=0.01)
    ctl = dict(tolerance=1e-12, relTol=0.0)
    for step in range(3):
        rp, rc = ref.step(nCorr=2, UControls=ctl, pControls=ctl)
        dp_, dc = dev.step(nCorr=2, UControls=ctl, pControls=ctl)
        assert len(dp_["U"]) == 3 and len(dp_["p"]) == len(rp["p"]) == 2
        assert all(p.converged for p in dp_["U"]) and all(p.converged for p in dp_["p"])
        for a, b in zip(dp_["p"], rp["p"]):
            assert abs(a.nIterations - b.nIterations) <= 2
        np.testing.assert_allclose(host(dev.U, 3), ref.U, rtol=0, atol=1e-9)
        np.testing.assert_allclose(host(dev.p), ref.p, rtol=0, atol=1e-9)
        np.testing.assert_allclose(host(dev.phi), ref.phi, rtol=0, atol=1e-11)
        assert dc[-1][0] < 1e-10 and rc[-1][0] < 1e-10
    # physics, on the device fields: the top layer follows the lid, mass is conserved
    cc = m.cell_centres()
    assert host(dev.U, 3)[cc[:, 1] > 1 - m.h, 0].mean() > 0.1
    assert float(dev.div(dev.phi, dev.bphi).abs().max()) < 1e-7
    dev.close()


@pytest.mark.parametrize("psolver,psecond", [("PCG", "preconditioner DIC;"),
                                             ("GAMG", "smoother GaussSeidel; nCellsInCoarsestLevel 10; mergeLevels 1;")])
def test_icofoam_case_directory_on_the_device(gpu, meshmod, orc, tmp_path, psolver, psecond):
    """icofoam.run_case on a written cavity case (polyMesh + dictionaries + 0/U, 0/p), device fields, against the
    oracle's steps; the pressure solver comes from system/fvSolution (PCG+DIC or GAMG+GaussSeidel)."""
    from oracle import piso_oracle as po
    from test_icofoam_case_cpu import write_cavity
    capi, ctx, torch = gpu
    ff = importlib.import_module("rapidcfd-dev_b200.foamfile")
    ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
    n = 8
    root = str(tmp_path / "cavity")
    m, dt = write_cavity(ff, meshmod, root, n, psolver, psecond, steps=2)
    lines = []
    case, hist = ico.run_case(capi, ctx, torch, root, log=lines.append)
    _, ref = po.cavity_from_hex(orc, meshmod, n, nu=0.01, deltaT=dt)
    ctl = dict(tolerance=1e-12, relTol=0.0)
    for _ in range(2):
        ref.step(UControls=ctl, pControls=ctl)
    np.testing.assert_allclose(host(case.U, 3), ref.U, rtol=0, atol=1e-8)
    np.testing.assert_allclose(host(case.p), ref.p, rtol=0, atol=1e-8)
    text = "\n".join(lines)
    assert text.count("Solving for p") == 4 and ("GAMG:" in text) == (psolver == "GAMG")
    assert all(c[0] < 1e-9 for _, cont in hist for c in cont)
    case.close()


@pytest.mark.parametrize("world", [2, 4])
def test_multi_gpu_fvm(world):
    """the glue over processor patches, one rank per GPU (needs >= 2 GPUs; skipped otherwise)"""
    import subprocess
    import sys
    import torch
    if os.environ.get("B200LDU_DRYRUN_ORACLE") == "1" or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + world),
           os.path.join(root, "tests", "multi_gpu_fvm_worker.py")]
    p = subprocess.run(cmd, cwd=root, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True, text=True,
                       timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-4000:]
    assert p.stdout.count("MULTI-GPU-FVM-OK") == 3 * world


@pytest.mark.parametrize("world", [2, 8])
def test_multi_gpu_icofoam(world):
    """the decomposed cavity, one rank per GPU, against the single-domain oracle (needs >= 2 GPUs; skipped otherwise)"""
    import subprocess
    import sys
    import torch
    if os.environ.get("B200LDU_DRYRUN_ORACLE") == "1" or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29800 + world),
           os.path.join(root, "tests", "multi_gpu_icofoam_worker.py")]
    p = subprocess.run(cmd, cwd=root, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True, text=True,
                       timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-4000:]
    assert p.stdout.count("MULTI-GPU-ICOFOAM-OK") == world


def test_simple_iterations_match_the_oracle(gpu, meshmod, orc):
    """SIMPLE (simpleFoam UEqn.H / pEqn.H: upwind convection, UEqn.relax, p.relax) sequenced over the C ABI against the
    numpy restatement of the same statements (oracle/piso_oracle.py simple_step)"""
    from oracle import piso_oracle as po
    capi, ctx, torch = gpu
    ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
    n = 8
    m, ref = po.cavity_from_hex(orc, meshmod, n, nu=0.05)
    m2, dev = ico.cavity(capi, ctx, torch, n, nu=0.05)
    ctl = dict(tolerance=1e-12, relTol=0.0)
    for it in range(4):
        rp, rc = ref.simple_step(UControls=ctl, pControls=ctl)
        dp_, dc = dev.simple_step(UControls=ctl, pControls=ctl)
        assert all(p.converged for p in dp_["U"]) and all(p.converged for p in dp_["p"])
        np.testing.assert_allclose(host(dev.U, 3), ref.U, rtol=0, atol=1e-9)
        np.testing.assert_allclose(host(dev.p), ref.p, rtol=0, atol=1e-9)
        np.testing.assert_allclose(host(dev.phi), ref.phi, rtol=0, atol=1e-11)
        assert abs(dc[1]) < 1e-10 and abs(rc[1]) < 1e-10
    dev.close()


def test_gpu_fvm_vs_golden(gpu, meshmod, orc):
    """the glue and two icoFoam steps against tests/golden/fvm_golden.npz (self-generated; make_fvm_golden.py)"""
    here = os.path.dirname(os.path.abspath(__file__))
    import sys
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_fvm_golden as mg
    capi, ctx, torch = gpu
    G = np.load(os.path.join(here, "golden", "fvm_golden.npz"))
    for nc in (1, 3):
        m, a, d, x, make_ = mg.fvm_case(meshmod, orc, nc)
        D = Dev(gpu, m, d, nc, x)
        assert np.array_equal(host(D.fv.A()), G[f"fvm{nc}.A"])
        assert np.array_equal(host(D.fv.H(), nc), G[f"fvm{nc}.H"])
        fi, fb, _ = D.fv.flux(D.nB)
        assert np.array_equal(host(fi, nc), G[f"fvm{nc}.flux"]) and np.array_equal(host(fb, nc), G[f"fvm{nc}.bflux"])
        if nc == 1:
            assert np.array_equal(host(D.fv.residual()), G["fvm1.residual"])
        D.fv.relax(0.7)
        assert np.array_equal(host(D.diag), G[f"fvm{nc}.relaxDiag"])
        assert np.array_equal(host(D.source, nc), G[f"fvm{nc}.relaxSource"])
        Z = Dev(gpu, m, d, nc, np.zeros((m.nCells, nc)), addr=D.addr)
        perfs = Z.fv.solve("PCG" if nc == 1 else "PBiCG", "DIC" if nc == 1 else "DILU", tolerance=1e-12, maxIter=500)
        assert all(abs(p.nIterations - k) <= 2 for p, k in zip(perfs, G[f"fvm{nc}.nIter"]))
        np.testing.assert_allclose(host(Z.psi, nc), G[f"fvm{nc}.psi"], rtol=0, atol=1e-9)
        Z.mat.close()
        D.close()
    ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
    _, dev = ico.cavity(capi, ctx, torch, mg.PISO_N)
    for _ in range(mg.PISO_STEPS):
        dev.step(UControls=mg.PISO_CTL, pControls=mg.PISO_CTL)
    np.testing.assert_allclose(host(dev.U, 3), G["piso.U"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(host(dev.p), G["piso.p"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(host(dev.phi), G["piso.phi"], rtol=0, atol=1e-11)
    dev.close()


def test_fvc_compositions_on_the_device(gpu, meshmod):
    """snGrad kernel, corrected snGrad and explicit Laplacian (rapidcfd-dev_b200/fvc.py) on a sheared mesh"""
    from test_fvc_cpu import run_sheared_case
    capi, ctx, torch = gpu
    run_sheared_case(meshmod, capi, ctx, torch)
