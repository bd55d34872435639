"""The device code of csrc/fvmatrix.cu and csrc/fieldops.cu, executed on the host (tests/host_kernels/harness.cpp
compiles the same *_kernels.cuh sources with the CUDA qualifiers defined away) and compared bit for bit with the
oracle.  This is how the fvMatrix glue kernels were checked before their first GPU run; the `-m gpu` tests of
tests/test_zzz_fvm_gpu.py hold the real launches against the same oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import fvm_oracle as fo
from test_oracle_core import _cyclic_case
from test_oracle_fvm import make, momentum_case, poisson_case

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_kernels", "harness.cpp")
OUT = os.path.join(HERE, "host_kernels", "_build", "libhostk.so")
CSRC = os.path.join(os.path.dirname(HERE), "rapidcfd-dev_b200", "csrc")


class HostCase(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nCells", "nFaces", "nB", "nC")] + \
               [(n, C.c_void_p) for n in ("l", "u", "ownerStart", "losortStart", "losort", "bStart", "bFaces", "bFaceCells",
                                          "cStart", "cFaces", "cFaceCells", "diag", "upper", "lower", "couInt", "couBou")]


@pytest.fixture(scope="module")
def hk():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("fvmatrix_kernels.cuh", "fieldops_kernels.cuh", "mules_kernels.cuh", "fv_kernels.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-Wall",
                               "-Wno-unused-function", "-Wno-unknown-pragmas", "-I", CSRC, "-shared", "-o", OUT, SRC])
    return C.CDLL(OUT)


def _csr(nCells, cells):
    start = np.zeros(nCells + 1, np.int32)
    np.add.at(start, np.asarray(cells, np.int64) + 1, 1)
    start = np.cumsum(start).astype(np.int32)
    faces = np.argsort(cells, kind="stable").astype(np.int32)
    return start, faces


class Host:
    """the arrays csrc/fvmatrix.cu reads, as the library builds them (b200ldu_fv_boundary_set, coupled_lists)"""

    def __init__(self, a, d, cfc=None, couInt=None, couBou=None):
        f64 = lambda x: None if x is None else np.ascontiguousarray(x, dtype=np.float64)
        i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
        self.keep = dict(l=i32(a.lower()), u=i32(a.upper()), ownerStart=i32(a.owner_start()), losortStart=i32(a.losort_start()),
                         losort=i32(a.losort()), bFaceCells=i32(d["bfc"]), diag=f64(d["diag"]), upper=f64(d["upper"]),
                         lower=f64(d["lower"] if d["lower"] is not None else d["upper"]))
        self.keep["bStart"], self.keep["bFaces"] = _csr(a.nCells, d["bfc"])
        cfc = np.zeros(0, np.int32) if cfc is None else i32(cfc)
        self.keep["cFaceCells"] = cfc
        self.keep["cStart"], self.keep["cFaces"] = _csr(a.nCells, cfc)
        self.keep["couInt"], self.keep["couBou"] = f64(np.zeros(1) if couInt is None else couInt), f64(np.zeros(1) if couBou is None else couBou)
        self.h = HostCase(a.nCells, a.nFaces, len(d["bfc"]), len(cfc),
                          **{k: v.ctypes.data for k, v in self.keep.items()})
        self.n = a.nCells

    def p(self):
        return C.byref(self.h)


def _d(x):
    return None if x is None else x.ctypes.data_as(C.c_void_p)


def run_all(hk, orc, a, d, nc, x, cfc=None, couInt=None, couBou=None):
    kw = {} if cfc is None else dict(couInt=couInt, couBou=couBou)
    ofm = make(orc, a, d, nc, x, **kw)
    H = Host(a, d, cfc, couInt, couBou)
    n, nB, nC = a.nCells, len(d["bfc"]), 0 if cfc is None else len(cfc)
    ic, bc = np.ascontiguousarray(d["ic"]), np.ascontiguousarray(d["bc"])
    psi, V = np.ascontiguousarray(x), np.ascontiguousarray(d["V"])
    source = np.ascontiguousarray(np.array(d["source"], float).reshape(n, nc))
    pnf = np.ascontiguousarray(ofm.patchNeighbourField()) if nC else None
    # A
    out = np.zeros(n)
    hk.hk_A(H.p(), nc, _d(ic), _d(V), _d(out))
    assert np.array_equal(out, ofm.A())
    # H
    out = np.zeros((n, nc))
    hk.hk_H(H.p(), nc, _d(psi), _d(source), _d(bc), _d(pnf), _d(V), _d(out))
    assert np.array_equal(out, ofm.H())
    # flux
    fi, fb, fc = np.zeros((a.nFaces, nc)), np.zeros((max(nB, 1), nc)), np.zeros((max(nC, 1), nc))
    hk.hk_flux(H.p(), nc, _d(psi), _d(ic), _d(bc), _d(pnf), _d(fi), _d(fb), _d(fc))
    oi, ob, oc = ofm.flux()
    assert np.array_equal(fi, oi) and np.array_equal(fb[:nB], ob) and np.array_equal(fc[:nC], oc)
    # boundary folding, in place
    for cmpt in list(range(nc)) + [-1]:
        got, ref = d["diag"].copy(), d["diag"].copy()
        hk.hk_boundary_diag(H.p(), nc, cmpt, _d(ic), _d(got), _d(got))
        ofm.addBoundaryDiag(ref, cmpt) if cmpt >= 0 else ofm.addCmptAvBoundaryDiag(ref)
        assert np.array_equal(got, ref)
    got = np.zeros(n)
    hk.hk_boundary_diag(H.p(), nc, 0, _d(ic), None, _d(got))             # diagIn NULL = zero
    ref = np.zeros(n)
    ofm.addBoundaryDiag(ref, 0)
    assert np.array_equal(got, ref)
    for couples in (False, True):
        got, ref = source.copy(), source.copy()
        hk.hk_boundary_source(H.p(), nc, _d(bc), _d(pnf) if couples else None, _d(got), _d(got))
        ofm.addBoundarySource(ref, couples)
        assert np.array_equal(got, ref)
    # the component loop of solveSegregated: source with everything in, then per component the coupled part out
    if nc == 3:
        tot = source.copy()
        ofm.addBoundarySource(tot, True)
        for k in range(3):
            got = np.zeros(n)
            hk.hk_component(H.p(), 3, k, 1, _d(pnf), _d(tot), _d(got))
            ref = np.ascontiguousarray(tot[:, k])
            if nC:
                np.subtract.at(ref, cfc, couBou * pnf[:, k])
            assert np.array_equal(got, ref)
            back = np.zeros((n, 3))
            hk.hk_set_component(H.p(), 3, k, _d(got), _d(back))
            assert np.array_equal(back[:, k], got) and not back[:, (k + 1) % 3].any()
    # residual's folded source
    if nc == 1:
        got = np.zeros(n)
        hk.hk_residual_source(H.p(), _d(ic), _d(psi), _d(source), _d(got))
        bd = np.zeros(n)
        ofm.addBoundaryDiag(bd, 0)
        assert np.array_equal(got, source[:, 0] - bd * psi[:, 0])
    # relax
    for alpha in (1.0, 0.6):
        orl = make(orc, a, d, nc, x, **kw)
        orl.relax(alpha)
        dg, sr = d["diag"].copy(), source.copy()
        hk.hk_relax(H.p(), nc, C.c_double(alpha), _d(psi), _d(ic), _d(dg), _d(sr))
        assert np.array_equal(dg, orl.diag) and np.array_equal(sr, orl.source)
    # setReference
    val = np.array([0.5, -1.25, 2.0][:nc])
    osr = make(orc, a, d, nc, x, **kw)
    osr.setReference(3, val)
    dg, sr = d["diag"].copy(), source.copy()
    hk.hk_set_reference(3, nc, _d(val), _d(dg), _d(sr))
    assert np.array_equal(dg, osr.diag) and np.array_equal(sr, osr.source)


@pytest.mark.parametrize("nc", [1, 3])
def test_fvmatrix_kernels_on_host_bit_exact(hk, meshmod, orc, nc):
    m, a, d = (poisson_case if nc == 1 else momentum_case)(meshmod, orc, (7, 6, 5))
    x = np.random.default_rng(2).uniform(-1, 1, (m.nCells, nc))
    run_all(hk, orc, a, d, nc, x)


@pytest.mark.parametrize("nc", [1, 3])
def test_fvmatrix_kernels_with_coupled_patches_on_host(hk, meshmod, orc, nc):
    m, c, ps, fc, nr, lo, hi = _cyclic_case(meshmod, "U")
    rng = np.random.default_rng(8)
    wall = np.concatenate([p.faceCells for p in m.wall_patches() if p.name not in ("xmin", "xmax")]).astype(np.int32)
    value = rng.uniform(-1, 1, (len(wall), nc))
    ic, bc = fo.fixedValue_laplacian_coeffs(np.full(len(wall), 0.01 * m.h * m.h), np.full(len(wall), 2.0 / m.h), value)
    diag = c["diag"].copy()
    np.subtract.at(diag, fc, c["int"])
    d = dict(diag=diag, upper=c["upper"], lower=c["lower"], source=rng.uniform(-1, 1, (m.nCells, nc)) * m.h ** 3,
             bfc=wall, ic=-ic, bc=-bc, V=m.volumes())
    a = orc.Addr(m.nCells, m.lower, m.upper, ps, fc, neighbRank=nr)
    x = rng.uniform(-1, 1, (m.nCells, nc))
    run_all(hk, orc, a, d, nc, x, cfc=fc, couInt=c["int"], couBou=c["bou"])


def test_fieldops_kernels_on_host(hk):
    rng = np.random.default_rng(0)
    n = 1000
    s, s2 = rng.uniform(-2, 2, n), rng.uniform(0.5, 2, n)
    v, v2 = rng.uniform(-2, 2, (n, 3)), rng.uniform(0.5, 2, (n, 3))
    ops = [np.add, np.subtract, np.multiply, np.divide, np.minimum, np.maximum]
    for op, f in enumerate(ops):
        for (A, nca), (B, ncb) in (((s, 1), (s2, 1)), ((v, 3), (v2, 3)), ((s, 1), (v2, 3)), ((v, 3), (s2, 1))):
            out = np.zeros((n, max(nca, ncb)))
            hk.hk_field_binary(op, C.c_longlong(n), nca, _d(np.ascontiguousarray(A)), ncb, _d(np.ascontiguousarray(B)), _d(out))
            ref = f(A.reshape(n, nca), B.reshape(n, ncb))
            assert np.array_equal(out, ref), (op, nca, ncb)
    sc = 0.37
    un = [lambda a: -a, np.abs, lambda a: sc * a, lambda a: sc / a, lambda a: a + sc, lambda a: sc - a,
          lambda a: np.minimum(a, sc), lambda a: np.maximum(a, sc), lambda a: a - sc, lambda a: a / sc, lambda a: a]
    flat = np.ascontiguousarray(v2.ravel())
    for op, f in enumerate(un):
        out = np.zeros(flat.size)
        hk.hk_field_unary(op, C.c_longlong(flat.size), C.c_double(sc), _d(flat), _d(out))
        assert np.array_equal(out, f(flat)), op
    out = np.zeros(n)
    hk.hk_field_dot3(C.c_longlong(n), _d(np.ascontiguousarray(v)), _d(np.ascontiguousarray(v2)), _d(out))
    assert np.array_equal(out, (v[:, 0] * v2[:, 0] + v[:, 1] * v2[:, 1]) + v[:, 2] * v2[:, 2])
    cells = rng.integers(0, n, 77).astype(np.int32)
    for nc, f in ((1, s), (3, v)):
        out = np.zeros((77, nc))
        hk.hk_field_gather(77, nc, _d(cells), _d(np.ascontiguousarray(f)), _d(out))
        assert np.array_equal(out, f.reshape(n, nc)[cells])


def test_sngrad_kernel_on_host(hk, meshmod, orc):
    m = meshmod.hex_mesh(5, 4, 3)
    a = orc.Addr(m.nCells, m.lower, m.upper)
    rng = np.random.default_rng(3)
    delta = rng.uniform(1, 2, m.nFaces)
    l, u = np.ascontiguousarray(m.lower, dtype=np.int32), np.ascontiguousarray(m.upper, dtype=np.int32)
    for nc in (1, 3):
        vf = rng.uniform(-1, 1, (m.nCells, nc))
        out = np.zeros((m.nFaces, nc))
        hk.hk_sngrad(m.nFaces, nc, _d(l), _d(u), _d(delta), _d(np.ascontiguousarray(vf)), _d(out))
        assert np.array_equal(out, np.asarray(orc.sngrad(a, delta, vf.ravel(), nc)).reshape(m.nFaces, nc))
        assert np.array_equal(out, delta[:, None] * (vf[m.upper] - vf[m.lower]))
