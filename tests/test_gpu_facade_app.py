"""The OpenFOAM-shaped C++ facade (foam/b200Foam.H) end to end: the mini-application reads
fvSolution-style dictionaries, selects solvers through lduMatrix::solver::New and prints
the reference's solverPerformance line."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "rapidcfd-dev_b200", "lib", "cavityPressureSolve")

DICTS = {
    "AINVPCG": "solvers { p { solver PCG; preconditioner DIC; tolerance 1e-07; relTol 0; } }",
    "diagonalPCG": "solvers { p { solver PCG; preconditioner { preconditioner diagonal; } tolerance 1e-07; } }",
    "GAMG": "solvers { p { solver GAMG; smoother GaussSeidel; tolerance 1e-07; relTol 0; nPreSweeps 0; "
            "nPostSweeps 2; cacheAgglomeration on; agglomerator faceAreaPair; nCellsInCoarsestLevel 10; "
            "mergeLevels 1; } }",
    "smoothSolver": "solvers { p { solver smoothSolver; smoother GaussSeidel; nSweeps 2; tolerance 1e-3; "
                    "maxIter 200; } } // comment",
}


@pytest.mark.parametrize("name", sorted(DICTS))
def test_app_solves(tmp_path, name):
    assert os.path.exists(APP), "run __graft_entry__.build()"
    f = tmp_path / "fvSolution"
    f.write_text(DICTS[name])
    p = subprocess.run([APP, "24", str(f)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    m = re.search(r"^(\w+):  Solving for p, Initial residual = ([\d.e+-]+), Final residual = ([\d.e+-]+), "
                  r"No Iterations (\d+)", p.stdout, re.M)
    assert m, p.stdout
    assert m.group(1) == name
    assert float(m.group(2)) == 1.0
    if name != "smoothSolver":
        assert float(m.group(3)) < 1e-7
        r = float(re.search(r"\|b - A p\|_1 / \|b\|_1 = ([\d.e+-]+)", p.stdout).group(1))
        assert r < 1e-5


def test_app_unknown_solver_lists_table(tmp_path):
    f = tmp_path / "fvSolution"
    f.write_text("solvers { p { solver PBiCG; preconditioner DILU; } }")  # asymmetric-only solver
    p = subprocess.run([APP, "8", str(f)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 2
    assert "Unknown symmetric matrix solver PBiCG" in p.stderr and "PCG" in p.stderr


def test_app_solver_type_registered_from_outside_the_facade(tmp_path):
    """run-time selection is a real registry: the application adds `loggedPCG` to the symmetric-matrix table with an
    addsymMatrixConstructorToTable object (as PCG.C:36-37 does in the reference) and the dictionary selects it"""
    f = tmp_path / "fvSolution"
    f.write_text("solvers { p { solver loggedPCG; preconditioner DIC; tolerance 1e-07; relTol 0; } }")
    p = subprocess.run([APP, "16", str(f)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert re.search(r"^loggedPCG: \d+ iterations", p.stdout, re.M) and "AINVPCG:  Solving for p" in p.stdout
    f.write_text("solvers { p { solver nonsense; } }")
    p = subprocess.run([APP, "8", str(f)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 2 and "loggedPCG" in p.stderr   # the valid-solver list now shows the added word
