"""ctypes access to oracle/_ref/libref_ldu.so -- the REFERENCE'S OWN lduMatrix sources
(lduMatrixATmul.C, lduAddressingFunctors.H, AINVPreconditionerF.H, JacobiSmootherF.H, ...) compiled
for the host against the shims in oracle/ref_harness/ (see harness.cpp).  TEST INFRASTRUCTURE ONLY:
used by tests/test_reference_functors.py to pin the oracle's row arithmetic to the code it restates.
Built by `make -C oracle ref` where /root/reference exists; elsewhere the prebuilt library is used."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libref_ldu.so")
_lib = None


def build():
    """(Re)build where the reference tree is present; returns the library path or None."""
    if os.path.isdir("/root/reference/src/OpenFOAM/matrices/lduMatrix"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    return _LIB if os.path.exists(_LIB) else None


def available():
    return build() is not None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref_ldu.so is not built (needs /root/reference)")
        _lib = C.CDLL(_LIB)
    return _lib


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _d(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def interface_update(nCells, faceCells, coeffs, pnf, result, negate=False):
    """coupledFvPatchField::updateInterfaceMatrix through the reference's matrixPatchOperation /
    matrixInterfaceFunctor: returns result with result[faceCells] -= coeffs*pnf (negate: +=)."""
    fc, c, v = _i(faceCells), _d(coeffs), _d(pnf)
    out = _d(result).copy()
    assert lib().ref_interface_update(int(nCells), len(fc), _p(fc), _p(c), _p(v), int(bool(negate)), _p(out)) == 0
    return out


class RefMatrix:
    """One single-domain LDU matrix handed to the reference code (no coupled interfaces)."""

    def __init__(self, nCells, lower, upper, ownerStart, losortStart, losort, diag, upperC, lowerC=None):
        self.n, self.nF = int(nCells), len(lower)
        self.keep = [_i(lower), _i(upper), _i(ownerStart), _i(losortStart), _i(losort), _d(diag), _d(upperC), _d(lowerC)]

    def _case(self):
        l, u, os_, ls, lo, dg, up, low = self.keep
        return [self.n, self.nF, _p(l), _p(u), _p(os_), _p(ls), _p(lo), _p(dg), _p(up), _p(low)]

    def op(self, which, favourSpeed=0, x=None, b=None):
        """which: amul | tmul | sumA | residual | H1 (lduMatrixATmul.C) | H | faceH (lduMatrixTemplates.C) |
        negSumDiag | sumDiag | sumMagOffDiag (lduMatrixOperations.C compositions; `b` = starting diagonal)."""
        code = {"amul": 0, "tmul": 1, "sumA": 2, "residual": 3, "H1": 4, "H": 5, "faceH": 6, "negSumDiag": 7,
                "sumDiag": 8, "sumMagOffDiag": 9}[which]
        out = np.zeros(self.nF if which == "faceH" else self.n)
        if code >= 7 and b is not None:
            out[:] = b
        x, b = _d(x), _d(b)
        rc = lib().ref_matrix_op(code, int(favourSpeed), *self._case(), _p(x), _p(b), _p(out))
        assert rc == 0
        return out

    def ainv(self, r, fast=False, transpose=False):
        out = np.zeros(self.n)
        r = _d(r)
        assert lib().ref_ainv(int(fast), int(transpose), *self._case(), _p(r), _p(out)) == 0
        return out

    def jacobi(self, psi, b, omega=0.9, fast=False):
        out = np.zeros(self.n)
        psi, b = _d(psi), _d(b)
        lib().ref_jacobi.argtypes = [C.c_int, C.c_double] + [C.c_int, C.c_int] + [C.c_void_p] * 11
        assert lib().ref_jacobi(int(fast), float(omega), *self._case(), _p(psi), _p(b), _p(out)) == 0
        return out


_LIB_GAMG = os.path.join(_HERE, "_ref", "libref_gamg.so")
_libg = None


def pair_agglomerate(nCells, lower, upper, faceWeights, forward=1):
    """The reference's pairGAMGAgglomeration::agglomerate (pairGAMGAgglomerate.C:135-313).
    Returns (map int32[nCells], nCoarseCells, forward flag after the call)."""
    global _libg
    if _libg is None:
        if not available() or not os.path.exists(_LIB_GAMG):
            raise RuntimeError("oracle/_ref/libref_gamg.so is not built (needs /root/reference)")
        _libg = C.CDLL(_LIB_GAMG)
    l, u, w = _i(lower), _i(upper), _d(faceWeights)
    fwd = C.c_int(int(forward))
    out = np.zeros(int(nCells), dtype=np.int32)
    nC = _libg.ref_pair_agglomerate(int(nCells), len(l), _p(l), _p(u), _p(w), C.byref(fwd), _p(out))
    return out, nC, fwd.value


_LIB_SOLVERS = os.path.join(_HERE, "_ref", "libref_solvers.so")
_libs = None


_LIB_SOLVERS_OMP = os.path.join(_HERE, "_ref", "libref_solvers_omp.so")
_libs_omp = None


def omp_available():
    """True when the all-core build of the reference solver loops exists (thrust host algorithms on OpenMP)."""
    return available() and os.path.exists(_LIB_SOLVERS_OMP)


_LIB_BINDING = os.path.join(_HERE, "_ref", "libref_binding.so")
_libs_binding = None


def binding_available():
    """oracle/_ref/libref_binding.so: the reference's solver sources + the product's reference-side binding
    (rapidcfd-dev_b200/foam/b200Solver.H), linked against libb200ldu.so"""
    return os.path.exists(_LIB_BINDING)


def solve(solver, precond, nCells, lower, upper, ownerStart, losortStart, losort, diag, upperC, lowerC, psi0, source,
          tolerance=1e-6, relTol=0.0, maxIter=1000, minIter=0, favourSpeed=0, nSweeps=1, omega=-1.0, omp=False,
          binding=False):
    """The reference's PCG::solve / PBiCG::solve / PBiCGStab::solve (PCG.C:69-208, PBiCG.C:68-246,
    PBiCGStab.C:66-300) with its own preconditioner classes, or smoothSolver::solve (smoothSolver.C:77-193,
    `precond` = smoother word, nSweeps, omega < 0 = not in the dictionary) with its JacobiSmoother.  Returns (psi, dict(initialResidual,
    finalResidual, nIterations, converged, singular, solverName))."""
    global _libs, _libs_omp, _libs_binding
    cur = _libs_binding if binding else (_libs_omp if omp else _libs)
    if cur is None:
        path = _LIB_BINDING if binding else (_LIB_SOLVERS_OMP if omp else _LIB_SOLVERS)
        if not os.path.exists(path) or not (binding or available()):
            raise RuntimeError(f"{path} is not built (needs /root/reference)")
        if binding:
            try:  # libb200ldu.so needs NCCL: let torch load its bundled (newer) libnccl first, as capi.lib() does
                import torch  # noqa: F401
            except ImportError:
                pass
        L = C.CDLL(path)
        L.ref_solve.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 8 + \
            [C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int,
             C.c_int, C.c_double]
        if binding:
            _libs_binding = L
        elif omp:
            _libs_omp = L
        else:
            _libs = L
    _L = _libs_binding if binding else (_libs_omp if omp else _libs)
    l, u, os_, ls, lo = _i(lower), _i(upper), _i(ownerStart), _i(losortStart), _i(losort)
    dg, up, low = _d(diag), _d(upperC), _d(lowerC)
    psi = _d(psi0).copy()
    src = _d(source)
    perf = np.zeros(5)
    name = C.create_string_buffer(512)
    rc = _L.ref_solve(solver.encode(), precond.encode(), int(favourSpeed), int(nCells), len(l), _p(l), _p(u), _p(os_),
                         _p(ls), _p(lo), _p(dg), _p(up), _p(low), float(tolerance), float(relTol), int(maxIter),
                         int(minIter), _p(psi), _p(src), _p(perf), name, 512, int(nSweeps), float(omega))
    if rc == -3:
        raise RuntimeError("FatalError: " + name.value.decode())
    if rc != 0:
        raise ValueError({-1: "unknown solver", -2: "unknown preconditioner"}.get(rc, rc))
    return psi, dict(initialResidual=perf[0], finalResidual=perf[1], nIterations=int(perf[2]), converged=bool(perf[3]),
                     singular=bool(perf[4]), solverName=name.value.decode())


_LIB_GAMGSOLVE = os.path.join(_HERE, "_ref", "libref_gamgsolve.so")
_libgs = None


def coarse_matrix(r, fr, flip, nCoarse, nCoarseFaces, diag, upper, lower):
    """Coarse-level coefficients by summation (GAMGSolverAgglomerateMatrix.C:37-322): restrict(diag), coarse
    faces = sums of their fine faces (upper/lower swapped where the face is flipped), faces collapsed into
    a coarse cell add 2*upper (symmetric) or upper+lower to its diagonal -- every sum in ascending fine
    index.  Plain numpy (np.add.at is sequential), used to feed the reference's V-cycle."""
    Dc = np.zeros(nCoarse)
    np.add.at(Dc, r, diag)
    Uc = np.zeros(nCoarseFaces)
    Lc = None if lower is None else np.zeros(nCoarseFaces)
    inside = fr < 0
    keep = ~inside
    if lower is None:
        np.add.at(Uc, fr[keep], upper[keep])
        np.add.at(Dc, -1 - fr[inside], 2 * upper[inside])
    else:
        fl = flip.astype(bool)
        np.add.at(Uc, fr[keep], np.where(fl, lower, upper)[keep])
        np.add.at(Lc, fr[keep], np.where(fl, upper, lower)[keep])
        np.add.at(Dc, -1 - fr[inside], (upper + lower)[inside])
    return Dc, Uc, Lc


def ldu_arrays(nCells, lower, upper):
    """ownerStart, losortStart, losort of an LDU addressing (lduAddressing.C:169-344: faces are sorted by
    owner already; losort = faces stably sorted by neighbour)."""
    l, u = _i(lower), _i(upper)
    n = int(nCells)
    ownerStart = np.concatenate([[0], np.cumsum(np.bincount(l, minlength=n))]).astype(np.int32)
    losortStart = np.concatenate([[0], np.cumsum(np.bincount(u, minlength=n))]).astype(np.int32)
    return ownerStart, losortStart, np.argsort(u, kind="stable").astype(np.int32)


def reference_hierarchy(nCells, lower, upper, faceWeights, nCellsInCoarsestLevel, diag, upperC, lowerC, forward=1,
                        maxLevels=50):
    """A GAMG hierarchy produced by REFERENCE CODE ONLY: pairGAMGAgglomeration::agglomerate for the maps,
    GAMGAgglomeration::agglomerateLduAddressing for coarse addressing / face maps, the reference's functors
    for the coarse coefficients; the level loop and the stop rule are pairGAMGAgglomerate.C:46-107 /
    GAMGAgglomeration.C:72-84 (continue while nCoarseCells >= nCellsInCoarsestLevel), the face weights are
    restricted by summation over the face map.  Returns a list of level dicts, finest first."""
    levels = [dict(nCells=int(nCells), lower=_i(lower), upper=_i(upper), diag=_d(diag), upperC=_d(upperC),
                   lowerC=_d(lowerC))]
    w = _d(faceWeights)
    fwd = forward
    while len(levels) < maxLevels:
        cur = levels[-1]
        rmap, nC, fwd = pair_agglomerate(cur["nCells"], cur["lower"], cur["upper"], w, fwd)
        if nC < nCellsInCoarsestLevel:
            break
        R = coarse_levels(cur["nCells"], cur["lower"], cur["upper"], rmap, nC, diag=cur["diag"], upperC=cur["upperC"],
                          lowerC=cur["lowerC"])
        cur["restrict"], cur["faceRestrict"], cur["flip"] = rmap, R["faceRestrict"], R["flip"]
        cw = np.zeros(len(R["coarseOwner"]))
        keep = R["faceRestrict"] >= 0
        np.add.at(cw, R["faceRestrict"][keep], w[keep])
        w = cw
        levels.append(dict(nCells=nC, lower=R["coarseOwner"].copy(), upper=R["coarseNeighbour"].copy(),
                           diag=R["coarseDiag"].copy(), upperC=R["coarseUpper"].copy(),
                           lowerC=None if R["coarseLower"] is None else R["coarseLower"].copy()))
    return levels, fwd


def oracle_hierarchy(g, addr, diag, upperC, lowerC):
    """The same level list built from the oracle's Gamg object `g` (maps + coarse addressing), coarse
    coefficients from coarse_matrix()."""
    levels = [dict(nCells=addr.nCells, lower=_i(addr.lower()), upper=_i(addr.upper()), diag=_d(diag), upperC=_d(upperC),
                   lowerC=_d(lowerC))]
    for k in range(g.nLevels):
        cur = levels[-1]
        r, fr, fl = g.restrict_addr(k), g.face_restrict_addr(k), g.face_flip(k)
        cur["restrict"] = _i(r)
        d, u, lo = coarse_matrix(r, fr, fl, g.ncells(k), g.nfaces(k), cur["diag"], cur["upperC"], cur["lowerC"])
        la = g.level_addr(k)
        levels.append(dict(nCells=g.ncells(k), lower=_i(la.lower()), upper=_i(la.upper()), diag=_d(d), upperC=_d(u),
                           lowerC=_d(lo)))
    return levels


def gamg_solve(g, addr, diag, upper, lower, psi0, source, **kw):
    """The reference's GAMGSolver::solve on the hierarchy of the oracle's Gamg object `g`."""
    return gamg_solve_levels(oracle_hierarchy(g, addr, diag, upper, lower), psi0, source, **kw)


def gamg_solve_levels(levels, psi0, source, tolerance=1e-6, relTol=0.0, maxIter=1000, minIter=0,
                      nPreSweeps=0, preSweepsLevelMultiplier=1, maxPreSweeps=4, nPostSweeps=2,
                      postSweepsLevelMultiplier=1, maxPostSweeps=4, nFinestSweeps=2, interpolateCorrection=0,
                      scaleCorrection=None, directSolveCoarsest=1, omega=-1.0, favourSpeed=0):
    """The reference's GAMGSolver::solve / Vcycle (GAMGSolverSolve.C) on a level list (finest first; every
    level but the last carries `restrict`).  Defaults are GAMGSolver.C:67-77's.  Returns (psi, dict) or raises
    NotImplementedError where the reference aborts."""
    global _libgs
    if _libgs is None:
        if not available() or not os.path.exists(_LIB_GAMGSOLVE):
            raise RuntimeError("oracle/_ref/libref_gamgsolve.so is not built (needs /root/reference)")
        _libgs = C.CDLL(_LIB_GAMGSOLVE)
    nL = len(levels) - 1
    lower = levels[0]["lowerC"]
    maps = [_i(lv["restrict"]) for lv in levels[:-1]]
    coeffs = [(lv["diag"], lv["upperC"], lv["lowerC"]) for lv in levels]
    derived = [ldu_arrays(lv["nCells"], lv["lower"], lv["upper"]) for lv in levels]
    keep = []

    def ptrs(arrs, ct):
        arr = (C.c_void_p * len(arrs))(*[(a.ctypes.data if a is not None else None) for a in arrs])
        keep.append((arrs, arr))
        return arr
    nC = _i([lv["nCells"] for lv in levels])
    nF = _i([len(lv["lower"]) for lv in levels])
    L = [lv["lower"] for lv in levels]
    U = [lv["upper"] for lv in levels]
    OS = [d[0] for d in derived]
    LS = [d[1] for d in derived]
    LO = [d[2] for d in derived]
    if scaleCorrection is None:
        scaleCorrection = 1 if lower is None else 0   # GAMGSolver.C:73: matrix.symmetric()
    ctl = _i([nPreSweeps, preSweepsLevelMultiplier, maxPreSweeps, nPostSweeps, postSweepsLevelMultiplier,
              maxPostSweeps, nFinestSweeps, interpolateCorrection, scaleCorrection, directSolveCoarsest, maxIter,
              minIter, favourSpeed])
    psi = _d(psi0).copy()
    src = _d(source)
    perf = np.zeros(5)
    _libgs.ref_gamg_solve.argtypes = [C.c_int] + [C.c_void_p] * 12 + [C.c_double, C.c_double, C.c_double] + [C.c_void_p] * 3
    rc = _libgs.ref_gamg_solve(nL, _p(nC), _p(nF), ptrs(L, 0), ptrs(U, 0), ptrs(OS, 0), ptrs(LS, 0), ptrs(LO, 0),
                               ptrs([c[0] for c in coeffs], 0), ptrs([c[1] for c in coeffs], 0),
                               ptrs([c[2] for c in coeffs], 0), ptrs(maps, 0), _p(ctl), float(tolerance), float(relTol),
                               float(omega), _p(psi), _p(src), _p(perf))
    if rc == -3:
        raise NotImplementedError("the reference aborts here: notImplemented(GAMGSolver::interpolate())")
    assert rc == 0
    return psi, dict(initialResidual=perf[0], finalResidual=perf[1], nIterations=int(perf[2]), converged=bool(perf[3]),
                     singular=bool(perf[4]))


_LIB_GAMGADDR = os.path.join(_HERE, "_ref", "libref_gamgaddr.so")
_libga = None


def coarse_levels(nCells, lower, upper, map0, nCoarse0, map1=None, nCoarse1=0, diag=None, upperC=None, lowerC=None):
    """The reference's GAMGAgglomeration::agglomerateLduAddressing (GAMGAgglomerateLduAddressing.C:245-603)
    for the restrict map `map0`; with `map1` (level 1 -> level 2) also the next level and then
    combineLevels(1) (:606-765).  Returns dict(restrict, faceRestrict, flip, coarseOwner, coarseNeighbour,
    nCoarseCells) describing level 0 afterwards; with `diag`/`upperC`/`lowerC` (single step) also the coarse
    coefficients assembled by GAMG::restrict / symAgglomerate / asymAgglomerate / diag*Agglomerate over the
    reference-built sorted addressing (GAMGAgglomerationTemplates.C:35-61, GAMGSolverAgglomerateMatrix.C:183-320)."""
    global _libga
    if _libga is None:
        if not available() or not os.path.exists(_LIB_GAMGADDR):
            raise RuntimeError("oracle/_ref/libref_gamgaddr.so is not built (needs /root/reference)")
        _libga = C.CDLL(_LIB_GAMGADDR)
    l, u, m0 = _i(lower), _i(upper), _i(map0)
    m1 = None if map1 is None else _i(map1)
    nF = len(l)
    r = np.zeros(int(nCells), np.int32)
    fr = np.zeros(max(nF, 1), np.int32)
    fl = np.zeros(max(nF, 1), np.uint8)
    co = np.zeros(max(nF, 1), np.int32)
    cn = np.zeros(max(nF, 1), np.int32)
    ncf = C.c_int(0)
    dg, up, low = _d(diag), _d(upperC), _d(lowerC)
    cD = np.zeros(max(int(nCoarse0), 1))
    cU, cL = np.zeros(max(nF, 1)), np.zeros(max(nF, 1))
    _libga.ref_coarse_levels.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_int] + \
        [C.c_void_p] * 12
    nc = _libga.ref_coarse_levels(2 if m1 is not None else 1, int(nCells), nF, _p(l), _p(u), _p(m0), int(nCoarse0),
                                  _p(m1), int(nCoarse1), _p(r), _p(fr), _p(fl), C.byref(ncf), _p(co), _p(cn),
                                  _p(dg), _p(up), _p(low), _p(cD), _p(cU), _p(cL))
    if nc < 0:
        raise RuntimeError("the reference code raised a FatalError")
    k = ncf.value
    out = dict(restrict=r, faceRestrict=fr[:nF], flip=fl[:nF], coarseOwner=co[:k], coarseNeighbour=cn[:k],
               nCoarseCells=nc)
    if diag is not None:   # coarse coefficients assembled by the reference's functors (single step)
        out.update(coarseDiag=cD[:nc], coarseUpper=cU[:k], coarseLower=None if lowerC is None else cL[:k])
    return out


_LIB_LDUADDR = os.path.join(_HERE, "_ref", "libref_lduaddr.so")
_libla = None


def ldu_addressing(nCells, lower, upper, patchStart=None, faceCells=None):
    """The reference's lduAddressing.C: ownerStart, losortStart, losort, ownerSort (thrust sorts / scans,
    :169-400), the per-patch sort addressing (:38-167) and band() (:453-498).  Returns a dict; the patch
    entries are lists (one array per patch)."""
    global _libla
    if _libla is None:
        if not available() or not os.path.exists(_LIB_LDUADDR):
            raise RuntimeError("oracle/_ref/libref_lduaddr.so is not built (needs /root/reference)")
        _libla = C.CDLL(_LIB_LDUADDR)
    l, u = _i(lower), _i(upper)
    n, nF = int(nCells), len(l)
    ps = _i(patchStart if patchStart is not None else [0])
    fc = _i(faceCells if faceCells is not None else [])
    nP, tot = len(ps) - 1, int(ps[-1])
    os_, ls = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.int32)
    lo, osrt = np.zeros(max(nF, 1), np.int32), np.zeros(max(nF, 1), np.int32)
    psa, psc, pss = (np.zeros(max(tot, 1), np.int32) for _ in range(3))
    nu = np.zeros(max(nP, 1), np.int32)
    band = np.zeros(2)
    rc = _libla.ref_ldu_addressing(n, nF, _p(l), _p(u), nP, _p(ps), _p(fc), _p(os_), _p(ls), _p(lo), _p(osrt), _p(psa),
                                   _p(psc), _p(pss), _p(nu), _p(band))
    if rc != 0:
        raise RuntimeError("the reference code raised a FatalError")
    out = dict(ownerStart=os_, losortStart=ls, losort=lo[:nF], ownerSort=osrt[:nF], bandwidth=int(band[0]),
               profile=band[1], patchSortAddr=[], patchSortCells=[], patchSortStart=[])
    for p in range(nP):
        s, e, k = int(ps[p]), int(ps[p + 1]), int(nu[p])
        out["patchSortAddr"].append(psa[s:e].copy())
        out["patchSortCells"].append(psc[s:s + k].copy())
        out["patchSortStart"].append(np.append(pss[s:s + k], e - s))
    return out


_LIB_FV = os.path.join(_HERE, "_ref", "libref_fv.so")
_libfv = None


def surface_integrate(nCells, lower, upper, ssf, bFaceCells, bssf, V, integrate=True):
    """The reference's fvc::surfaceIntegrate(ivf, ssf) for a scalar surface field
    (fvcSurfaceIntegrate.C:41-205): per cell, owner faces added, neighbour faces subtracted, then the
    boundary faces (handed over as one patch, see harness_fv.cpp), then the division by the volumes.
    integrate=False: fvc::surfaceSum (:261-352) -- all faces added, no division."""
    global _libfv
    if _libfv is None:
        if not available() or not os.path.exists(_LIB_FV):
            raise RuntimeError("oracle/_ref/libref_fv.so is not built (needs /root/reference)")
        _libfv = C.CDLL(_LIB_FV)
    l, u = _i(lower), _i(upper)
    os_, ls, lo = ldu_arrays(nCells, l, u)
    s, bs, v = _d(ssf), _d(bssf), _d(V)
    bfc = _i(bFaceCells)
    out = np.zeros(int(nCells))
    _libfv.ref_surface_integrate(int(bool(integrate)), int(nCells), len(l), _p(l), _p(u), _p(_i(os_)), _p(_i(ls)), _p(_i(lo)), _p(s),
                                 len(bfc), _p(bfc), _p(bs), _p(v), _p(out))
    return out


def surface_integrate_vec(nCells, lower, upper, ssf, bFaceCells, bssf, V, integrate=True):
    """fvc::surfaceIntegrate / surfaceSum of a VECTOR surface field ((faces, 3) arrays) -> (nCells, 3)"""
    surface_integrate(1, [], [], [], [], [], [1.0])  # loads the library
    l, u = _i(lower), _i(upper)
    os_, ls, lo = ldu_arrays(nCells, l, u)
    s, bs, v, bfc = _d(np.ravel(ssf)), _d(np.ravel(bssf)), _d(V), _i(bFaceCells)
    out = np.zeros(3 * int(nCells))
    _libfv.ref_surface_integrate_vec(int(bool(integrate)), int(nCells), len(l), _p(l), _p(u), _p(_i(os_)), _p(_i(ls)),
                                     _p(_i(lo)), _p(s), len(bfc), _p(bfc), _p(bs), _p(v), _p(out))
    return out.reshape(-1, 3)


def gauss_gradf_vec(nCells, lower, upper, Sf, ssf, bFaceCells, bSf, bssf, V):
    """fv::gaussGrad<vector>::gradf: the tensor field T_ij = d_i u_j, (nCells, 3, 3)"""
    surface_integrate(1, [], [], [], [], [], [1.0])
    l, u = _i(lower), _i(upper)
    os_, ls, lo = ldu_arrays(nCells, l, u)
    A, s, bA, bs, v, bfc = _d(np.ravel(Sf)), _d(np.ravel(ssf)), _d(np.ravel(bSf)), _d(np.ravel(bssf)), _d(V), _i(bFaceCells)
    out = np.zeros(9 * int(nCells))
    _libfv.ref_gauss_gradf_vec(int(nCells), len(l), _p(l), _p(u), _p(_i(os_)), _p(_i(ls)), _p(_i(lo)), _p(A), _p(s), len(bfc),
                               _p(bfc), _p(bA), _p(bs), _p(v), _p(out))
    return out.reshape(-1, 3, 3)


def gauss_gradf(nCells, lower, upper, Sf, ssf, bFaceCells, bSf, bssf, V):
    """The reference's fv::gaussGrad<scalar>::gradf (gaussGrad.C:34-243): per cell the sum of Sf*ssf over the
    owner faces, minus the neighbour faces, plus the boundary faces, divided by the volume.  Returns (nCells, 3)."""
    surface_integrate(1, [], [], [], [], [], [1.0])  # loads the library
    l, u = _i(lower), _i(upper)
    os_, ls, lo = ldu_arrays(nCells, l, u)
    A, s, bA, bs, v = _d(np.ravel(Sf)), _d(ssf), _d(np.ravel(bSf)), _d(bssf), _d(V)
    bfc = _i(bFaceCells)
    out = np.zeros(3 * int(nCells))
    _libfv.ref_gauss_gradf(int(nCells), len(l), _p(l), _p(u), _p(_i(os_)), _p(_i(ls)), _p(_i(lo)), _p(A), _p(s),
                           len(bfc), _p(bfc), _p(bA), _p(bs), _p(v), _p(out))
    return out.reshape(-1, 3)


_LIB_GAMGIFACE = os.path.join(_HERE, "_ref", "libref_gamgiface.so")
_libgi = None


class InterfaceAgglomeration:
    """The reference's multi-rank coarse-level construction (GAMGAgglomerateLduAddressing.C with its interface
    branch, GAMGInterface.C, processorGAMGInterface.C): all ranks of a decomposed case in this process.
    ranks: list of dicts(nCells, lower, upper, patchStart, faceCells, neighbRank)."""

    def __init__(self, ranks):
        global _libgi
        if _libgi is None:
            if not available() or not os.path.exists(_LIB_GAMGIFACE):
                raise RuntimeError("oracle/_ref/libref_gamgiface.so is not built (needs /root/reference)")
            _libgi = C.CDLL(_LIB_GAMGIFACE)
        self.nRanks = len(ranks)
        self.h = _libgi.ref_ia_create(self.nRanks)
        self.fine = []          # per level, per rank: (nCells, nFaces, patch sizes) of the FINE side
        lev0 = []
        for r, d in enumerate(ranks):
            l, u, ps, fc, nb = _i(d["lower"]), _i(d["upper"]), _i(d["patchStart"]), _i(d["faceCells"]), _i(d["neighbRank"])
            _libgi.ref_ia_set_rank(self.h, r, int(d["nCells"]), len(l), _p(l), _p(u), len(ps) - 1, _p(ps), _p(fc), _p(nb))
            lev0.append((int(d["nCells"]), len(l), np.diff(ps).astype(int)))
        self.fine.append(lev0)

    def agglomerate(self, level, maps, nCoarse):
        """maps[r]: restrict map of rank r from level to level+1.  Returns per rank a dict with the coarse
        addressing, the face restrict / flip maps, coarse patchStart / faceCells and the patch-face restrict map."""
        for r in range(self.nRanks):
            m = _i(maps[r])
            _libgi.ref_ia_set_map(self.h, level, r, _p(m), len(m), int(nCoarse[r]))
        if _libgi.ref_ia_agglomerate(self.h, level) != 0:
            raise RuntimeError("the reference code raised a FatalError")
        return self._collect(level, level)

    def combine(self, level):
        """combineLevels(level): folds level into level-1 on every rank; returns the new description of level-1."""
        if _libgi.ref_ia_combine(self.h, level) != 0:
            raise RuntimeError("the reference code raised a FatalError")
        self.fine.pop()
        return self._collect(level - 1, level - 1)

    def _collect(self, level, fineLevel):
        out, nxt = [], []
        for r in range(self.nRanks):
            nFineCells, nFineFaces, finePatch = self.fine[fineLevel][r]
            sz = np.zeros(3 + len(finePatch) + 1, np.int32)
            _libgi.ref_ia_sizes(self.h, level, r, _p(sz))
            nC, nCF, nP = int(sz[0]), int(sz[1]), int(sz[2])
            cps = np.concatenate([[0], np.cumsum(sz[3:3 + nP])]).astype(np.int32)
            own, nei = np.zeros(max(nCF, 1), np.int32), np.zeros(max(nCF, 1), np.int32)
            fr, fl = np.zeros(max(nFineFaces, 1), np.int32), np.zeros(max(nFineFaces, 1), np.uint8)
            cfc = np.zeros(max(int(cps[-1]), 1), np.int32)
            pfr = np.zeros(max(int(finePatch.sum()), 1), np.int32)
            _libgi.ref_ia_get(self.h, level, r, _p(own), _p(nei), _p(fr), fl.ctypes.data_as(C.c_void_p), _p(cfc), _p(pfr))
            out.append(dict(nCells=nC, lower=own[:nCF], upper=nei[:nCF], faceRestrict=fr[:nFineFaces],
                            faceFlip=fl[:nFineFaces], patchStart=cps, faceCells=cfc[:int(cps[-1])],
                            patchFaceRestrict=pfr[:int(finePatch.sum())], finePatchSizes=finePatch))
            nxt.append((nC, nCF, np.diff(cps).astype(int)))
        if len(self.fine) == fineLevel + 1:
            self.fine.append(nxt)
        else:
            self.fine[fineLevel + 1] = nxt
        return out

    def agglomerate_coeffs(self, level, r, patch, fine):
        f = _d(fine)
        out = np.zeros(max(len(f), 1))
        n = _libgi.ref_ia_coeffs(self.h, level, r, int(patch), _p(f), len(f), _p(out))
        if n < 0:
            raise RuntimeError("the reference code raised a FatalError")
        return out[:n]

    def __del__(self):
        if _libgi is not None and getattr(self, "h", None) is not None:
            _libgi.ref_ia_destroy(self.h)
            self.h = None


_LIB_FVM = os.path.join(_HERE, "_ref", "libref_fvm.so")
_libfvm = None
FVM_OPS = dict(addBoundaryDiag=0, addCmptAvBoundaryDiag=1, addBoundarySource=2, setReference=3, relax=4, D=5, A=6, flux=7,
               H=8, residual=9, solveSegregated=10)


def fvm(op, nCells, lower, upper, patches, V, psi, diag, upperC, lowerC, source, iarg=0, darg=0.0, x=None, nc=1):
    """The reference's fvMatrix<scalar> (fvMatrix.C, fvScalarMatrix.C compiled for the host) on one matrix.
    patches: list of dict(faceCells, ic, bc, coupled=False, pnf=None) in mesh order.  Returns the op's outputs:
    addBoundary*: the updated x; setReference / relax: (diag, source); D, A, H, residual: the field; flux: (internal faces,
    boundary faces flat); solveSegregated: (diagonal seen by the solver, source seen by the solver, diagonal afterwards).
    nc = 3: fvMatrix<vector> (fields (n, 3), patch ic / bc / pnf (faces, 3)); setReference, flux and residual are scalar
    only; solveSegregated returns one (nCells) block per component in the first two outputs."""
    global _libfvm
    if _libfvm is None:
        if not available() or not os.path.exists(_LIB_FVM):
            raise RuntimeError("oracle/_ref/libref_fvm.so is not built (needs /root/reference)")
        _libfvm = C.CDLL(_LIB_FVM)
    l, u = _i(lower), _i(upper)
    n, nF = int(nCells), len(l)
    os_, ls, lo = ldu_arrays(n, l, u)
    ps = np.zeros(len(patches) + 1, np.int32)
    for k, p in enumerate(patches):
        ps[k + 1] = ps[k] + len(p["faceCells"])
    cat = lambda key, dt, dflt=None: (np.concatenate([np.asarray(p.get(key) if p.get(key) is not None else dflt(p), dt).ravel()
                                                      for p in patches]) if patches else np.zeros(0, dt))
    fc = _i(cat("faceCells", np.int32))
    ic, bc = _d(cat("ic", float)), _d(cat("bc", float))
    pnf = _d(cat("pnf", float, lambda p: np.zeros(len(p["faceCells"]) * nc)))
    coupled = _i([1 if p.get("coupled") else 0 for p in patches] or [0])
    tot = int(ps[-1])
    o1, o2, o3 = np.zeros(max(n, nF, 1) * 3), np.zeros(max(n, tot, 1) * 3), np.zeros(max(n, 1))
    xin = _d(np.zeros(n * nc) if x is None else np.ravel(x))
    lowp = None if lowerC is None else _p(_d(lowerC))
    d = [_d(np.ravel(a)) for a in (V, psi, diag, upperC, source)]
    rc = (_libfvm.ref_fvm if nc == 1 else _libfvm.ref_fvm_vec)(FVM_OPS[op], n, nF, _p(l), _p(u), _p(_i(os_)), _p(_i(ls)), _p(_i(lo)), len(patches), _p(ps), _p(fc),
                         _p(coupled), _p(pnf), _p(d[0]), _p(d[1]), _p(d[2]), _p(d[3]), lowp, _p(d[4]), _p(ic), _p(bc),
                         int(iarg), C.c_double(darg), _p(xin), _p(o1), _p(o2), _p(o3))
    if rc != 0:
        raise RuntimeError(f"the reference code raised a FatalError (rc {rc})")
    vec = lambda a: a[:n * nc].reshape(n, nc).copy() if nc > 1 else a[:n].copy()
    if op in ("addBoundaryDiag", "addCmptAvBoundaryDiag", "D", "A", "residual"):
        return o1[:n].copy()
    if op in ("addBoundarySource", "H"):
        return vec(o1)
    if op in ("setReference", "relax"):
        return o1[:n].copy(), vec(o2)
    if op == "flux":
        return o1[:nF].copy(), o2[:tot].copy()
    if nc == 1:
        return o1[:n].copy(), o2[:n].copy(), o3[:n].copy()
    return o1[:n * nc].reshape(nc, n).copy(), o2[:n * nc].reshape(nc, n).copy(), o3[:n].copy()


_LIB_PF = os.path.join(_HERE, "_ref", "libref_procfield.so")
_libpf = None


def processor_interface_update(ranks, commsType="nonBlocking", negate=False, nPoll=0, gpuDirect=False, gamgLevel=False):
    """The reference's finest-level coupled-interface update over all ranks of a decomposed case (in one process):
    lduMatrix::initMatrixInterfaces on every rank, then lduMatrix::updateMatrixInterfaces
    (lduMatrixUpdateMatrixInterfaces.C, processorFvPatchScalarField.C, matrixPatchOperation / matrixInterfaceFunctor).
    ranks: list of dict(nCells, patchStart, faceCells, neighbRank, coeffs, psi, result).  Returns the updated results.
    gamgLevel: the interfaces are processorGAMGInterfaceField objects (processorGAMGInterfaceField.C:94-248,
    GAMGUpdateInterfaceMatrix) as on the coarse GAMG levels instead of processorFvPatchField<scalar>."""
    global _libpf
    if _libpf is None:
        if not available() or not os.path.exists(_LIB_PF):
            raise RuntimeError("oracle/_ref/libref_procfield.so is not built (needs /root/reference)")
        _libpf = C.CDLL(_LIB_PF)
    _libpf.ref_pf_reset(len(ranks))
    for r, d in enumerate(ranks):
        ps, fc, nr = _i(d["patchStart"]), _i(d["faceCells"]), _i(d["neighbRank"])
        co, psi, res = _d(d["coeffs"]), _d(d["psi"]), _d(d["result"])
        _libpf.ref_pf_set_rank(r, int(d["nCells"]), len(ps) - 1, _p(ps), _p(fc), _p(nr), _p(co), _p(psi), _p(res),
                               int(bool(gamgLevel)))
    if _libpf.ref_pf_update({"blocking": 0, "nonBlocking": 2}[commsType], int(bool(negate)), int(nPoll), int(bool(gpuDirect))) != 0:
        raise RuntimeError("the reference code raised a FatalError")
    out = []
    for r, d in enumerate(ranks):
        res = np.zeros(int(d["nCells"]))
        _libpf.ref_pf_get(r, _p(res))
        out.append(res)
    return out


def fvm_fill(which, nc, nCells, lower, upper, patches, a, b):
    """Coefficient fills through the reference's schemes: which = "div": gaussConvectionScheme::fvmDiv(faceFlux = b, vf) with
    interpolation weights a; which = "laplacian": gaussLaplacianScheme::fvmLaplacianUncorrected(gammaMagSf = a, deltaCoeffs = b,
    vf).  patches: list of dict(faceCells, kind ("fixedValue" | "zeroGradient" | "coupled"), value (faces, nc) for fixedValue,
    delta = patch deltaCoeffs, a, b = the two surface fields on the patch faces).  Returns dict(lower (div only), upper, diag,
    ic, bc) with ic / bc flat over the patches, (faces, nc)."""
    fvm("D", 1, [], [], [], [1.0], [0.0], [1.0], [], None, [0.0])   # loads the library
    l, u = _i(lower), _i(upper)
    n, nF = int(nCells), len(l)
    os_, ls, lo = ldu_arrays(n, l, u)
    ps = np.zeros(len(patches) + 1, np.int32)
    for k, p in enumerate(patches):
        ps[k + 1] = ps[k] + len(p["faceCells"])
    tot = int(ps[-1])
    kinds = _i([{"fixedValue": 0, "zeroGradient": 1, "coupled": 2}[p["kind"]] for p in patches] or [0])
    cat = lambda f: _d(np.concatenate([np.ravel(f(p)) for p in patches]) if patches else np.zeros(1))
    fc = _i(np.concatenate([p["faceCells"] for p in patches]) if patches else [0])
    pvalue = cat(lambda p: np.asarray(p.get("value", np.zeros((len(p["faceCells"]), nc))), float))
    pdelta = cat(lambda p: np.asarray(p.get("delta", np.zeros(len(p["faceCells"]))), float))
    pa, pb = cat(lambda p: p["a"]), cat(lambda p: p["b"])
    A, B = _d(a), _d(b)
    low, upp, dg = np.zeros(max(nF, 1)), np.zeros(max(nF, 1)), np.zeros(n)
    ic, bc = np.zeros(max(tot * nc, 1)), np.zeros(max(tot * nc, 1))
    rc = _libfvm.ref_fvm_fill({"div": 0, "laplacian": 1}[which], int(nc), n, nF, _p(l), _p(u), _p(_i(os_)), _p(_i(ls)), _p(_i(lo)),
                              len(patches), _p(ps), _p(fc), _p(kinds), _p(pvalue), _p(pdelta), _p(A), _p(B), _p(pa), _p(pb),
                              _p(low), _p(upp), _p(dg), _p(ic), _p(bc))
    if rc != 0:
        raise RuntimeError("the reference code raised a FatalError")
    out = dict(upper=upp[:nF], diag=dg, ic=ic[:tot * nc].reshape(tot, nc), bc=bc[:tot * nc].reshape(tot, nc))
    if which == "div":
        out["lower"] = low[:nF]
    return out


def euler_ddt(nCells, lower, upper, patches, deltaT, V, U0, phi0, Sf, w):
    """EulerDdtScheme<vector>::fvmDdt(U) and ::fvcDdtPhiCorr(U, phi) (with ddtScheme::fvcDdtPhiCoeff) run from the reference's
    sources on a field whose old-time level is U0 and a flux whose old-time level is phi0.  patches: list of dict(faceCells,
    fixesValue (bool), value (faces, 3) = U0 on the patch, phi0 (faces), Sf (faces, 3)).  Returns dict(diag, source (n, 3),
    ddtCorr (nF), bddtCorr (flat over the patches))."""
    fvm("D", 1, [], [], [], [1.0], [0.0], [1.0], [], None, [0.0])   # loads the library
    l, u = _i(lower), _i(upper)
    n, nF = int(nCells), len(l)
    os_, ls, lo = ldu_arrays(n, l, u)
    ps = np.zeros(len(patches) + 1, np.int32)
    for k, p in enumerate(patches):
        ps[k + 1] = ps[k] + len(p["faceCells"])
    tot = int(ps[-1])
    cat = lambda f, w_: _d(np.concatenate([np.ravel(f(p)) for p in patches]) if patches else np.zeros(w_))
    fc = _i(np.concatenate([p["faceCells"] for p in patches]) if patches else [0])
    fixes = _i([1 if p["fixesValue"] else 0 for p in patches] or [0])
    bU, bphi, bSf = cat(lambda p: p["value"], 3), cat(lambda p: p["phi0"], 1), cat(lambda p: p["Sf"], 3)
    diag, source = np.zeros(n), np.zeros(n * 3)
    corr, bcorr = np.zeros(max(nF, 1)), np.zeros(max(tot, 1))
    V_, U0_, phi0_, Sf_, w_ = _d(V), _d(np.ravel(U0)), _d(phi0), _d(np.ravel(Sf)), _d(w)
    _libfvm.ref_ddt.restype = C.c_int
    rc = _libfvm.ref_ddt(n, nF, _p(l), _p(u), _p(_i(os_)), _p(_i(ls)), _p(_i(lo)), len(patches), _p(ps), _p(fc), _p(fixes),
                         C.c_double(float(deltaT)), _p(V_), _p(U0_), _p(bU), _p(phi0_), _p(bphi), _p(Sf_), _p(bSf), _p(w_),
                         _p(diag), _p(source), _p(corr), _p(bcorr))
    if rc != 0:
        raise RuntimeError("the reference code raised a FatalError")
    return dict(diag=diag, source=source.reshape(n, 3), ddtCorr=corr[:nF], bddtCorr=bcorr[:tot])


def interpolate_linear(nc, nCells, lower, upper, patches, w, vf):
    """surfaceInterpolationScheme<Type>::interpolate(vf) of the reference with linear weights w.  patches: list of
    dict(faceCells, coupled (bool), w (faces), value (faces, nc) = the patch field, pnf (faces, nc) = patchNeighbourField of a
    coupled patch).  Returns (internal face values (nF[, nc]), patch face values flat over the patches (tot[, nc]))."""
    fvm("D", 1, [], [], [], [1.0], [0.0], [1.0], [], None, [0.0])   # loads the library
    l, u = _i(lower), _i(upper)
    n, nF = int(nCells), len(l)
    os_, ls, lo = ldu_arrays(n, l, u)
    ps = np.zeros(len(patches) + 1, np.int32)
    for k, p in enumerate(patches):
        ps[k + 1] = ps[k] + len(p["faceCells"])
    tot = int(ps[-1])
    cat = lambda f: _d(np.concatenate([np.ravel(f(p)) for p in patches]) if patches else np.zeros(nc))
    fc = _i(np.concatenate([p["faceCells"] for p in patches]) if patches else [0])
    cpl = _i([1 if p["coupled"] else 0 for p in patches] or [0])
    zero = lambda p: np.zeros((len(p["faceCells"]), nc))
    pw, bvf, pnf = cat(lambda p: p["w"]), cat(lambda p: p.get("value", zero(p))), cat(lambda p: p.get("pnf", zero(p)))
    out, bout = np.zeros(max(nF * nc, 1)), np.zeros(max(tot * nc, 1))
    _libfvm.ref_interpolate.restype = C.c_int
    rc = _libfvm.ref_interpolate(int(nc), n, nF, _p(l), _p(u), _p(_i(os_)), _p(_i(ls)), _p(_i(lo)), len(patches), _p(ps), _p(fc),
                                 _p(cpl), _p(_d(w)), _p(pw), _p(_d(np.ravel(vf))), _p(bvf), _p(pnf), _p(out), _p(bout))
    if rc != 0:
        raise RuntimeError("the reference code raised a FatalError")
    shape = (lambda k: (k,)) if nc == 1 else (lambda k: (k, nc))
    return out[:nF * nc].reshape(shape(nF)), bout[:tot * nc].reshape(shape(tot))


_LIB_LDUOPS = os.path.join(_HERE, "_ref", "libref_lduops.so")
_liblduops = None


def ldu_combine(nCells, lower, upper, A, op1, B, op2=0, Cm=None):
    """lduMatrix::operator+= / operator-= of the reference (lduMatrixOperations.C:235-397): A (op1) B (op2) Cm with op = +1 / -1;
    each matrix a dict with any of diag, upper, lower (absent = the reference matrix has no such array).  Returns the dict of the
    arrays the result holds."""
    global _liblduops
    if _liblduops is None:
        if not os.path.exists(_LIB_LDUOPS):
            raise RuntimeError("oracle/_ref/libref_lduops.so is missing: run `make -C oracle ref` where /root/reference exists")
        _liblduops = C.CDLL(_LIB_LDUOPS)
        _liblduops.ref_ldu_combine.restype = C.c_int
    l, u = _i(lower), _i(upper)
    n, nF = int(nCells), len(l)
    keep = []

    def arr(m, k):
        if m is None or m.get(k) is None:
            return None
        keep.append(_d(m[k]))
        return _p(keep[-1])
    Cm = Cm or {}
    dO, uO, lO, has = np.zeros(n), np.zeros(max(nF, 1)), np.zeros(max(nF, 1)), np.zeros(3, np.int32)
    rc = _liblduops.ref_ldu_combine(n, nF, _p(l), _p(u), arr(A, "diag"), arr(A, "upper"), arr(A, "lower"), int(op1), arr(B, "diag"),
                                    arr(B, "upper"), arr(B, "lower"), int(op2), arr(Cm, "diag"), arr(Cm, "upper"), arr(Cm, "lower"),
                                    _p(dO), _p(uO), _p(lO), _p(has))
    if rc != 0:
        raise RuntimeError("the reference code raised a FatalError")
    out = {}
    if has[0]:
        out["diag"] = dO
    if has[1]:
        out["upper"] = uO[:nF]
    if has[2]:
        out["lower"] = lO[:nF]
    return out


def fvm_assemble(nCells, lower, upper, patches, V, A, B, Cm, su):
    """((A + B) - Cm) == su through the reference's fvMatrix operators (fvMatrix.C operator+ / operator- / operator== and
    operator+= / -=) for vector matrices over one field: A = dict(diag, source (n, 3)) diagonal, B = dict(diag, upper, lower, ic,
    bc) asymmetric, Cm = dict(diag, upper, ic, bc) symmetric; ic / bc flat over the patches (tot, 3); patches: list of face-cell
    arrays.  Returns dict(diag, upper, lower, source, ic, bc, kind)."""
    fvm("D", 1, [], [], [], [1.0], [0.0], [1.0], [], None, [0.0])   # loads the library
    l, u = _i(lower), _i(upper)
    n, nF = int(nCells), len(l)
    os_, ls, lo = ldu_arrays(n, l, u)
    ps = np.zeros(len(patches) + 1, np.int32)
    for k, fcs in enumerate(patches):
        ps[k + 1] = ps[k] + len(fcs)
    tot = int(ps[-1])
    fc = _i(np.concatenate(patches) if patches else [0])
    flat = lambda x: _d(np.ravel(x))
    keep = [flat(A["diag"]), flat(A["source"]), flat(B["diag"]), flat(B["upper"]), flat(B["lower"]), flat(B["ic"]), flat(B["bc"]),
            flat(Cm["diag"]), flat(Cm["upper"]), flat(Cm["ic"]), flat(Cm["bc"]), flat(su)]
    dg, up, low = np.zeros(n), np.zeros(max(nF, 1)), np.zeros(max(nF, 1))
    src, ic, bc = np.zeros(n * 3), np.zeros(max(tot * 3, 1)), np.zeros(max(tot * 3, 1))
    _libfvm.ref_fvm_assemble.restype = C.c_int
    rc = _libfvm.ref_fvm_assemble(n, nF, _p(l), _p(u), _p(_i(os_)), _p(_i(ls)), _p(_i(lo)), len(patches), _p(ps), _p(fc), _p(_d(V)),
                                  *[_p(k) for k in keep], _p(dg), _p(up), _p(low), _p(src), _p(ic), _p(bc))
    if rc < 0:
        raise RuntimeError("the reference code raised a FatalError")
    return dict(diag=dg, upper=up[:nF], lower=low[:nF], source=src.reshape(n, 3), ic=ic[:tot * 3].reshape(tot, 3),
                bc=bc[:tot * 3].reshape(tot, 3), kind={0: "diagonal", 1: "symmetric", 2: "asymmetric"}[rc])

