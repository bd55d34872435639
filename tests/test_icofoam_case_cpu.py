"""SURVEY.md section 8(f) ranks 2 + 3 together, host side: `icofoam.run_case` reads an OpenFOAM case directory
(polyMesh, transportProperties, controlDict, fvSolution, 0/U, 0/p), runs the time loop and writes the result.
The device calls go to the oracle-backed stand-in (tests/oracle_backend.py), so this checks the case reading, the
control flow and the log, not the CUDA path (tests/test_zzz_fvm_gpu.py holds that against the same oracle)."""
import importlib
import os

import numpy as np
import pytest

import oracle_backend
from oracle import piso_oracle as po

CONTROL = """FoamFile { version 2.0; format ascii; class dictionary; object controlDict; }
application icoFoam; startFrom startTime; startTime 0; stopAt endTime; endTime %r; deltaT %r;
writeControl timeStep; writeInterval 20;
"""
FVSOLUTION = """FoamFile { version 2.0; format ascii; class dictionary; object fvSolution; }
solvers
{
    p { solver %s; %s tolerance 1e-12; relTol 0; }
    U { solver PBiCG; preconditioner DILU; tolerance 1e-12; relTol 0; }
}
PISO { nCorrectors 2; nNonOrthogonalCorrectors 0; pRefCell 0; pRefValue 0; }
"""
TRANSPORT = """FoamFile { version 2.0; format ascii; class dictionary; object transportProperties; }
nu nu [0 2 -1 0 0 0 0] 0.01;
"""


def write_cavity(ff, meshmod, root, n, psolver="PCG", psecond="preconditioner DIC;", steps=2):
    m = meshmod.hex_mesh(n)
    pm = ff.from_hex_mesh(m)
    os.makedirs(os.path.join(root, "constant", "polyMesh"))
    os.makedirs(os.path.join(root, "system"))
    os.makedirs(os.path.join(root, "0"))
    ff.write_poly_mesh(os.path.join(root, "constant", "polyMesh"), pm)
    dt = 0.5 * m.h
    open(os.path.join(root, "system", "controlDict"), "w").write(CONTROL % (steps * dt, dt))
    open(os.path.join(root, "system", "fvSolution"), "w").write(FVSOLUTION % (psolver, psecond))
    open(os.path.join(root, "constant", "transportProperties"), "w").write(TRANSPORT)
    Ubf = {p.name: {"type": "fixedValue", "value": np.array([1.0, 0, 0]) if p.name == "movingWall" else np.zeros(3)}
           for p in pm.patches}
    pbf = {p.name: {"type": "zeroGradient"} for p in pm.patches}
    ff.write_field(os.path.join(root, "0", "U"), "volVectorField", [0, 1, -1, 0, 0, 0, 0], np.zeros(3), Ubf)
    ff.write_field(os.path.join(root, "0", "p"), "volScalarField", [0, 2, -2, 0, 0, 0, 0], 0.0, pbf)
    return m, dt


@pytest.mark.parametrize("psolver,psecond", [("PCG", "preconditioner DIC;"), ("GAMG", "smoother GaussSeidel; nCellsInCoarsestLevel 10; mergeLevels 1; agglomerator faceAreaPair;")])
def test_run_case_reads_the_case_and_matches_the_oracle_step(meshmod, orc, tmp_path, psolver, psecond):
    ff = importlib.import_module("rapidcfd-dev_b200.foamfile")
    ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
    n = 6
    root = str(tmp_path / "cavity")
    m, dt = write_cavity(ff, meshmod, root, n, psolver, psecond)
    capi, ctx, torch = oracle_backend.fixture()
    lines = []
    case, hist = ico.run_case(capi, ctx, torch, root, log=lines.append)
    assert len(hist) == 2
    # the same two steps straight on the oracle with the analytic hex geometry
    _, ref = po.cavity_from_hex(orc, meshmod, n, nu=0.01, deltaT=dt)
    ctl = dict(tolerance=1e-12, relTol=0.0)
    for _ in range(2):
        ref.step(UControls=ctl, pControls=ctl)
    U = case.U.numpy().reshape(-1, 3)
    np.testing.assert_allclose(U, ref.U, rtol=0, atol=1e-8)
    np.testing.assert_allclose(case.p.numpy(), ref.p, rtol=0, atol=1e-8)
    # the log carries the reference's lines: one per velocity component, two pressure solves and a continuity line per step
    text = "\n".join(lines)
    assert text.count("Solving for Ux") == 2 and text.count("Solving for p") == 4
    assert text.count("time step continuity errors") == 4 and "Time = " in text and lines[-1].startswith("End")
    name = "GAMG" if psolver == "GAMG" else "AINVPCG"
    assert f"{name}:  Solving for p, Initial residual" in text and "AINVPBiCG:  Solving for Ux" in text
    # the last time directory holds U and p as field files that read back
    tdir = os.path.join(root, ico._time_name(2 * dt))
    Uf = ff.read_field(os.path.join(tdir, "U"))
    pf = ff.read_field(os.path.join(tdir, "p"))
    np.testing.assert_allclose(Uf["internalField"], U, rtol=1e-15)
    np.testing.assert_allclose(pf["internalField"], case.p.numpy(), rtol=1e-15)
    assert str(Uf["boundaryField"]["movingWall"]["type"]) == "fixedValue"


def test_run_case_refuses_what_the_step_does_not_cover(meshmod, tmp_path):
    ff = importlib.import_module("rapidcfd-dev_b200.foamfile")
    ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
    root = str(tmp_path / "c")
    write_cavity(ff, meshmod, root, 4)
    pm = ff.read_poly_mesh(os.path.join(root, "constant", "polyMesh"))
    pbf = {p.name: {"type": "zeroGradient"} for p in pm.patches}
    pbf["movingWall"] = {"type": "fixedValue", "value": 0.0}
    ff.write_field(os.path.join(root, "0", "p"), "volScalarField", [0, 2, -2, 0, 0, 0, 0], 0.0, pbf)
    capi, ctx, torch = oracle_backend.fixture()
    with pytest.raises(NotImplementedError, match="movingWall"):
        ico.run_case(capi, ctx, torch, root, log=lambda s: None)


@pytest.mark.parametrize("nR", [2, 4])
def test_decomposed_device_step_sequencing(meshmod, orc, nR):
    """The multi-rank branch of icofoam.IcoFoam (processor patches: coupled interpolation, interface coefficients,
    patchNeighbourField exchanges, the zero-padded boundary list for the glue) through the stand-in, one thread per
    rank, against the single-domain oracle run."""
    import dist_helpers as dh
    ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
    n = 8
    ctl = dict(tolerance=1e-12, relTol=0.0)
    _, ref = po.cavity_from_hex(orc, meshmod, n)
    for _ in range(2):
        ref.step(UControls=ctl, pControls=ctl)
    ex = dh.ThreadExchange(nR)

    def rank_fn(r):
        m0 = meshmod.decompose(n, nR, r)
        comm = ex.comm(orc, r, m0, n ** 3)
        capi, _, torch = oracle_backend.fixture()
        ctx = oracle_backend._Ctx(comm)

        def allsum(v):
            out = np.array(v, float)
            orc.lib().orc_comm_sum(comm.ptr(), orc._d(out), len(out))
            return out
        m, case = ico.cavity_rank(capi, ctx, torch, n, nR, r, allsum)
        for _ in range(2):
            perfs, cont = case.step(UControls=ctl, pControls=ctl)
        return m.cellGlobal, case.U.numpy().reshape(-1, 3), case.p.numpy(), cont
    res = dh.run_threads(nR, rank_fn)
    for cg, U, p, cont in res:
        np.testing.assert_allclose(U, ref.U[cg], rtol=0, atol=1e-8)
        np.testing.assert_allclose(p, ref.p[cg], rtol=0, atol=1e-8)
        assert cont[-1][0] < 1e-10 and cont == res[0][3]


def test_run_case_on_processor_directories(meshmod, orc, tmp_path):
    """decomposePar layout: processorN/constant/polyMesh (+ cellProcAddressing), processorN/0/{U,p}, shared system/ and
    constant/transportProperties -- every rank reads its directory, exchanges cell centres for the coupled weights,
    runs the step over its processor patches and writes its time directory; the fields agree with the single domain."""
    import types
    import dist_helpers as dh
    ff = importlib.import_module("rapidcfd-dev_b200.foamfile")
    ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
    n, nR = 8, 4
    root = str(tmp_path / "cavity")
    m, dt = write_cavity(ff, meshmod, root, n, steps=2)
    pm = ff.read_poly_mesh(os.path.join(root, "constant", "polyMesh"))
    cellToProc = np.zeros(m.nCells, int)
    for r in range(nR):
        cellToProc[meshmod.decompose(n, nR, r).cellGlobal] = r
    parts = ff.decompose_poly_mesh(pm, cellToProc)
    for r, (sub, cells, faceG) in enumerate(parts):
        pdir = os.path.join(root, f"processor{r}")
        os.makedirs(os.path.join(pdir, "constant", "polyMesh"))
        os.makedirs(os.path.join(pdir, "0"))
        ff.write_poly_mesh(os.path.join(pdir, "constant", "polyMesh"), sub)
        ff.write_list(os.path.join(pdir, "constant", "polyMesh", "cellProcAddressing"), "label", cells.astype(np.int32))
        Ubf, pbf = {}, {}
        for p in sub.patches:
            if p.type == "processor":
                Ubf[p.name], pbf[p.name] = {"type": "processor"}, {"type": "processor"}
            else:
                Ubf[p.name] = {"type": "fixedValue", "value": np.array([1.0, 0, 0]) if p.name == "movingWall" else np.zeros(3)}
                pbf[p.name] = {"type": "zeroGradient"}
        ff.write_field(os.path.join(pdir, "0", "U"), "volVectorField", [0, 1, -1, 0, 0, 0, 0], np.zeros(3), Ubf)
        ff.write_field(os.path.join(pdir, "0", "p"), "volScalarField", [0, 2, -2, 0, 0, 0, 0], 0.0, pbf)
    _, ref = po.cavity_from_hex(orc, meshmod, n, nu=0.01, deltaT=dt)
    ctl = dict(tolerance=1e-12, relTol=0.0)
    for _ in range(2):
        ref.step(UControls=ctl, pControls=ctl)
    ex = dh.ThreadExchange(nR)

    def rank_fn(r):
        sub = parts[r][0]
        shim = types.SimpleNamespace(coupled_patches=lambda: [types.SimpleNamespace(neighbRank=p.neighbProcNo)
                                                              for p in sub.patches if p.type == "processor"])
        comm = ex.comm(orc, r, shim, n ** 3)
        capi, _, torch = oracle_backend.fixture()
        ctx = oracle_backend._Ctx(comm)

        def allsum(v):
            out = np.array(v, float)
            orc.lib().orc_comm_sum(comm.ptr(), orc._d(out), len(out))
            return out
        lines = []
        case, hist = ico.run_case(capi, ctx, torch, root, log=lines.append, rank=r, allsum=allsum)
        return case.U.numpy().reshape(-1, 3), case.p.numpy(), case.pRefCell, "\\n".join(lines)
    res = dh.run_threads(nR, rank_fn)
    assert sorted(r[2] for r in res)[:-1] == [-1] * (nR - 1) and max(r[2] for r in res) >= 0   # one rank owns the reference cell
    for r, (U, p, _, text) in enumerate(res):
        cells = parts[r][1]
        np.testing.assert_allclose(U, ref.U[cells], rtol=0, atol=1e-8)
        np.testing.assert_allclose(p, ref.p[cells], rtol=0, atol=1e-8)
        assert text.count("time step continuity errors") == 4
        tdir = os.path.join(root, f"processor{r}", ico._time_name(2 * dt))
        np.testing.assert_allclose(ff.read_field(os.path.join(tdir, "p"))["internalField"], p, rtol=1e-15)


FVSCHEMES = """FoamFile { version 2.0; format ascii; class dictionary; object fvSchemes; }
ddtSchemes { default Euler; }
gradSchemes { default Gauss linear; grad(p) Gauss linear; }
divSchemes { default none; div(phi,U) Gauss %s; }
laplacianSchemes { default none; laplacian(nu,U) Gauss linear orthogonal; laplacian((1|A(U)),p) Gauss linear orthogonal; }
interpolationSchemes { default linear; interpolate(HbyA) linear; }
snGradSchemes { default orthogonal; }
fluxRequired { default no; p; }
"""


def test_run_case_honours_fvSchemes(meshmod, orc, tmp_path):
    """div(phi,U) Gauss upwind from system/fvSchemes reaches the momentum matrix (keywords with parentheses are single words,
    as ISstream reads them); a scheme the step does not implement, or a missing entry under `default none`, is refused"""
    ff = importlib.import_module("rapidcfd-dev_b200.foamfile")
    ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
    n = 6
    root = str(tmp_path / "cavity")
    m, dt = write_cavity(ff, meshmod, root, n)
    schemes = os.path.join(root, "system", "fvSchemes")
    open(schemes, "w").write(FVSCHEMES % "upwind")
    capi, ctx, torch = oracle_backend.fixture()
    case, hist = ico.run_case(capi, ctx, torch, root, log=lambda *_: None, write=False)
    ctl = dict(tolerance=1e-12, relTol=0.0)
    refs = {}
    for scheme in ("upwind", "linear"):
        _, ref = po.cavity_from_hex(orc, meshmod, n, nu=0.01, deltaT=dt)
        for _ in range(2):
            ref.step(UControls=ctl, pControls=ctl, divScheme=scheme)
        refs[scheme] = ref.U.copy()
    U = case.U.numpy().reshape(-1, 3)
    np.testing.assert_allclose(U, refs["upwind"], rtol=0, atol=1e-8)
    assert np.abs(refs["upwind"] - refs["linear"]).max() > 1e-4          # the scheme is visible in the result
    open(schemes, "w").write(FVSCHEMES % "limitedLinearV 1")
    with pytest.raises(ValueError, match="Unknown discretisation scheme limitedLinearV"):
        ico.run_case(capi, ctx, torch, root, log=lambda *_: None, write=False)
    open(schemes, "w").write(FVSCHEMES % "vanLeer")
    with pytest.raises(NotImplementedError, match="scalar fields only"):
        ico.run_case(capi, ctx, torch, root, log=lambda *_: None, write=False)
    open(schemes, "w").write((FVSCHEMES % "linear").replace("div(phi,U) Gauss linear;", ""))
    with pytest.raises(KeyError, match=r"keyword div\(phi,U\) is undefined in dictionary"):
        ico.run_case(capi, ctx, torch, root, log=lambda *_: None, write=False)
    open(schemes, "w").write((FVSCHEMES % "linear").replace("default Euler", "default CrankNicolson 0.9"))
    with pytest.raises(NotImplementedError, match="ddt"):
        ico.run_case(capi, ctx, torch, root, log=lambda *_: None, write=False)
