"""Generates tests/golden/fvm_golden.npz -- SELF-GENERATED fixtures for the fvMatrix glue (row a17) and the icoFoam
step (section 8(f) rank 2).  Like ldu_golden.npz they do NOT pin the oracle to the reference (fvMatrix.C and the
application do not compile against a shim): they freeze the outputs of oracle/fvm_oracle.py and oracle/piso_oracle.py on
seeded inputs so that `pytest -m "not gpu"` notices drift bit for bit and `pytest -m gpu` compares the CUDA path with
vectors that exist independently of the oracle build on the GPU box.

    python tests/golden/make_fvm_golden.py      # rewrites fvm_golden.npz
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DIMS = (7, 6, 5)
PISO_N, PISO_STEPS = 6, 2
PISO_CTL = dict(tolerance=1e-12, relTol=0.0)


def fvm_case(meshmod, orc, nc):
    import test_oracle_fvm as tf
    m, a, d = (tf.poisson_case if nc == 1 else tf.momentum_case)(meshmod, orc, DIMS)
    x = np.random.default_rng(2).uniform(-1, 1, (m.nCells, nc))
    return m, a, d, x, tf.make


def generate(meshmod, orc):
    from oracle import piso_oracle as po
    out = {}
    for nc in (1, 3):
        m, a, d, x, make = fvm_case(meshmod, orc, nc)
        f = make(orc, a, d, nc, x)
        out[f"fvm{nc}.A"], out[f"fvm{nc}.H"] = f.A(), f.H()
        out[f"fvm{nc}.flux"], out[f"fvm{nc}.bflux"], _ = f.flux()
        if nc == 1:
            out["fvm1.residual"] = f.residual()
        r = make(orc, a, d, nc, x)
        r.relax(0.7)
        out[f"fvm{nc}.relaxDiag"], out[f"fvm{nc}.relaxSource"] = r.diag, r.source
        psi, perfs, _ = make(orc, a, d, nc).solve("PCG" if nc == 1 else "PBiCG", "DIC" if nc == 1 else "DILU",
                                                  tolerance=1e-12, maxIter=500)
        out[f"fvm{nc}.psi"] = psi
        out[f"fvm{nc}.nIter"] = np.array([p.nIterations for p in perfs])
    _, case = po.cavity_from_hex(orc, meshmod, PISO_N)
    for _ in range(PISO_STEPS):
        case.step(UControls=PISO_CTL, pControls=PISO_CTL)
    out["piso.U"], out["piso.p"], out["piso.phi"] = case.U, case.p, case.phi
    return out


if __name__ == "__main__":
    meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
    from oracle import ldu_oracle as orc
    orc.build()
    data = generate(meshmod, orc)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fvm_golden.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path), "bytes,", len(data), "arrays")
