"""CPU restatement (numpy) of the reference's explicit MULES -- TEST INFRASTRUCTURE (the checker of b200ldu_mules_limiter and of
rapidcfd-dev_b200/mules.py; never imported by the product).

  MULES::limiter        FV/fvMatrices/solvers/MULES/MULESTemplates.C:381-745 (functors :143-377, MULESFunctors.H)
  MULES::limit          :748-813  (phiBD = upwind flux, phiCorr = phiPsi - phiBD, phiPsi = phiBD + lambda*phiCorr)
  MULES::explicitSolve  :36-78    (psi = (rho0*psi0*rDeltaT + Su - surfaceIntegrate(phiPsi))/(rho*rDeltaT - Sp))
Static mesh, non-coupled boundary patches.  Per cell the faces are visited in the reference's order -- owner faces, neighbour
faces in losort order, boundary faces patch by patch -- and every operator is one fp64 rounding.  rho / rho0 None:
geometricOneField; Sp / Su None: zeroField (their operators return the other operand unchanged, one.H / zero.H).
Pinned: tests/test_mules_cpu.py runs the reference's own MULESTemplates.C (oracle/_ref/libref_mules.so) beside it."""
import ctypes as C
import os

import numpy as np

SMALL, VSMALL = 1e-15, 1e-300
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libref_mules.so")
_lib = None


class Incidence:
    """per cell: its faces in the reference's visiting order, padded.  face index into [internal faces, boundary faces];
    kind 0 owner, 1 neighbour, 2 boundary"""

    def __init__(self, nCells, lower, upper, bFaceCells):
        lower, upper, bfc = (np.asarray(a, np.int64) for a in (lower, upper, bFaceCells))
        nF = len(lower)
        cells = np.concatenate([lower, upper[np.argsort(upper, kind="stable")], bfc])
        faces = np.concatenate([np.arange(nF), np.argsort(upper, kind="stable"), nF + np.arange(len(bfc))])
        kinds = np.concatenate([np.zeros(nF, np.int8), np.ones(nF, np.int8), np.full(len(bfc), 2, np.int8)])
        order = np.lexsort((np.arange(len(cells)), kinds, cells))     # by cell, then kind, then the order within the kind
        cells, faces, kinds = cells[order], faces[order], kinds[order]
        count = np.bincount(cells, minlength=nCells)
        start = np.concatenate([[0], np.cumsum(count)])
        slot = np.arange(len(cells)) - start[cells]
        self.width = int(count.max()) if len(cells) else 0
        self.face = np.zeros((nCells, self.width), np.int64)
        self.kind = np.zeros((nCells, self.width), np.int8)
        self.valid = np.zeros((nCells, self.width), bool)
        self.face[cells, slot], self.kind[cells, slot], self.valid[cells, slot] = faces, kinds, True
        self.nF, self.lower, self.upper, self.bfc = nF, lower, upper, bfc

    def other(self):
        """the cell across each incident face (-1 on boundary faces)"""
        f = np.minimum(self.face, max(self.nF - 1, 0))
        return np.where(self.kind == 0, self.upper[f], np.where(self.kind == 1, self.lower[f], -1))


def _a(x):
    return None if x is None else np.asarray(x, np.float64)


def limiter_steps(nCells, lower, upper, bFaceCells, V, rDeltaT, psi, psi0, psiB, phiBD, phiBDB, phiCorr, phiCorrB, psiMax, psiMin,
                  nLimiterIter=3, rho=None, rho0=None, Sp=None, Su=None, lambda0=None, lambdaB0=None, nCoupled=0, corr=False,
                  extremaCoeff=0.0):
    """corr=True: MULES::limiterCorr (CMULESTemplates.C:375-704) -- the same sweeps around other budgets: no bounded flux (phiBD is
    ignored, phiBDB holds the boundary values of the TOTAL flux phi for the outflow test of patchLambdaPfCMULESFunctor), the extrema
    widened by extremaCoeff*(psiMax - psiMin), the current psi and rho in place of the old-time ones (psi0, rho0 are ignored).
    MULES::limiter as a generator: after every sweep it yields the limiters of the trailing nCoupled boundary faces (the coupled
    patch faces, for which psiB holds patchNeighbourField()) and is sent the other side's values -- syncTools::syncFaceList with
    minOp (MULESTemplates.C:743).  Returns (lambda, lambdaB) through StopIteration."""
    inc = Incidence(nCells, lower, upper, bFaceCells)
    psi, psi0, psiB, V = _a(psi), _a(psi0), _a(psiB), _a(V)
    bdAll, corrAll = np.concatenate([_a(phiBD), _a(phiBDB)]), np.concatenate([_a(phiCorr), _a(phiCorrB)])
    nF, nB = inc.nF, len(inc.bfc)
    other = inc.other()
    psiMaxn, psiMinn = np.full(nCells, float(psiMin)), np.full(nCells, float(psiMax))
    sumBD, sumPhip, mSumPhim = np.zeros(nCells), np.full(nCells, VSMALL), np.full(nCells, VSMALL)
    for k in range(inc.width):
        f, kind, ok = inc.face[:, k], inc.kind[:, k], inc.valid[:, k]
        pn = np.where(kind == 2, psiB[np.clip(f - nF, 0, max(nB - 1, 0))] if nB else 0.0, psi[np.maximum(other[:, k], 0)])
        psiMaxn = np.where(ok, np.maximum(psiMaxn, pn), psiMaxn)
        psiMinn = np.where(ok, np.minimum(psiMinn, pn), psiMinn)
        ownerLike = kind != 1
        if not corr:
            sumBD = np.where(ok, np.where(ownerLike, sumBD + bdAll[f], sumBD - bdAll[f]), sumBD)
        pc = corrAll[f]
        toP = (pc > 0) == ownerLike            # owner & positive, or neighbour & not positive -> sumPhip
        sumPhip = np.where(ok & toP, np.where(ownerLike, sumPhip + pc, sumPhip - pc), sumPhip)
        mSumPhim = np.where(ok & ~toP, np.where(ownerLike, mSumPhim - pc, mSumPhim + pc), mSumPhim)
    if corr:
        e = extremaCoeff * (psiMax - psiMin)
        psiMaxn, psiMinn = np.minimum(psiMaxn + e, psiMax), np.maximum(psiMinn - e, psiMin)
    else:
        psiMaxn, psiMinn = np.minimum(psiMaxn, psiMax), np.maximum(psiMinn, psiMin)
    a = rDeltaT if rho is None else _a(rho) * rDeltaT
    if Sp is not None:
        a = a - _a(Sp)
    if corr:                                        # rho*psi*rDeltaT (CMULESTemplates.C:508, :517)
        b = (psi if rho is None else _a(rho) * psi) * rDeltaT
    else:
        r0 = rho0 if rho0 is not None else rho      # rho.oldTime(): (rho0*rDeltaT)*psi0 (MULESTemplates.C:543, :552)
        b = (rDeltaT if r0 is None else _a(r0) * rDeltaT) * psi0
    up = a * psiMaxn
    if Su is not None:
        up = up - _a(Su)
    lo = -(a * psiMinn) if Su is None else _a(Su) - a * psiMinn
    if corr:
        psiMaxn, psiMinn = V * (up - b), V * (lo + b)
    else:
        psiMaxn, psiMinn = V * (up - b) + sumBD, V * (lo + b) - sumBD
    lamAll = np.ones(nF + nB)
    if lambda0 is not None:
        lamAll[:nF] = lambda0
    if lambdaB0 is not None and nB:
        lamAll[nF:] = lambdaB0
    for _ in range(nLimiterIter):
        slp, mslm = np.zeros(nCells), np.zeros(nCells)
        for k in range(inc.width):
            f, kind, ok = inc.face[:, k], inc.kind[:, k], inc.valid[:, k]
            lp = lamAll[f] * corrAll[f]
            ownerLike = kind != 1
            toP = (lp > 0) == ownerLike
            slp = np.where(ok & toP, np.where(ownerLike, slp + lp, slp - lp), slp)
            mslm = np.where(ok & ~toP, np.where(ownerLike, mslm - lp, mslm + lp), mslm)
        lambdam = np.maximum(np.minimum((slp + psiMaxn) / (mSumPhim - SMALL), 1.0), 0.0)
        lambdap = np.maximum(np.minimum((mslm + psiMinn) / (sumPhip + SMALL), 1.0), 0.0)
        o, n = inc.lower, inc.upper
        pc = corrAll[:nF]
        lamAll[:nF] = np.where(pc > 0, np.minimum(lamAll[:nF], np.minimum(lambdap[o], lambdam[n])),
                               np.minimum(lamAll[:nF], np.minimum(lambdam[o], lambdap[n])))
        if nB:
            pcB, c = corrAll[nF:], inc.bfc
            lim = np.where(pcB > 0, np.minimum(lamAll[nF:], lambdap[c]), np.minimum(lamAll[nF:], lambdam[c]))
            # non-coupled faces, outflow only: patchLambdaPfMULESFunctor (phiBD + phiCorr), patchLambdaPfCMULESFunctor (phi)
            outflow = (bdAll[nF:] if corr else bdAll[nF:] + pcB) > SMALL * SMALL
            outflow[nB - nCoupled:] = True                              # coupledPatchLambdaPfMULESFunctor: every face
            lamAll[nF:] = np.where(outflow, lim, lamAll[nF:])
        if nCoupled:
            theirs = yield lamAll[nF + nB - nCoupled:].copy()
            lamAll[nF + nB - nCoupled:] = np.minimum(lamAll[nF + nB - nCoupled:], theirs)
    return lamAll[:nF].copy(), lamAll[nF:].copy()


def limiter(*args, **kw):
    """single domain (no coupled faces): MULES::limiter -> (lambda, lambdaB)"""
    assert not kw.get("nCoupled")
    g = limiter_steps(*args, **kw)
    try:
        next(g)
    except StopIteration as e:
        return e.value
    raise AssertionError("unreachable: no coupled faces, nothing to synchronise")


def limiter_ranks(cases, exchange):
    """the ranks of a decomposed case in lockstep.  cases: one dict of limiter_steps keyword arguments per rank (nCoupled > 0, psiB
    of the coupled faces = the neighbour cells' psi); exchange(list of per-rank coupled arrays) -> list of the arrays each rank
    receives (the same faces seen from the other side).  Returns [(lambda, lambdaB)] per rank."""
    gens = [limiter_steps(**c) for c in cases]
    out = [None] * len(gens)
    try:
        mine = [next(g) for g in gens]
    except StopIteration:                      # nLimiterIter = 0: every generator returns at once
        return [limiter_steps_result(c) for c in cases]
    while True:
        theirs = exchange(mine)
        nxt, done = [], 0
        for r, g in enumerate(gens):
            try:
                nxt.append(g.send(theirs[r]))
            except StopIteration as e:
                out[r] = e.value
                done += 1
        if done:
            assert done == len(gens)
            return out
        mine = nxt


def limiter_steps_result(c):
    g = limiter_steps(**c)
    try:
        next(g)
    except StopIteration as e:
        return e.value
    raise AssertionError("expected no sweep")


def upwind_flux(lower, upper, phi, phiB, psi, psiB):
    """upwind<scalar>::flux: faceFlux*interpolate(psi) with the weights pos(faceFlux) (upwind.H:86-103,
    surfaceInterpolationScheme.C:263-321 one-weight form, :176-184); boundary faces: faceFlux*psi_b"""
    phi, psi = _a(phi), _a(psi)
    w = np.where(phi >= 0, 1.0, 0.0)
    sf = w * (psi[lower] - psi[upper]) + psi[upper]
    return phi * sf, _a(phiB) * _a(psiB)


def limit(nCells, lower, upper, bFaceCells, V, rDeltaT, psi, psi0, psiB, phi, phiB, phiPsi, phiPsiB, psiMax, psiMin,
          nLimiterIter=3, rho=None, rho0=None, Sp=None, Su=None):
    phiBD, phiBDB = upwind_flux(lower, upper, phi, phiB, psi, psiB)
    phiCorr, phiCorrB = _a(phiPsi) - phiBD, _a(phiPsiB) - phiBDB
    lam, lamB = limiter(nCells, lower, upper, bFaceCells, V, rDeltaT, psi, psi0, psiB, phiBD, phiBDB, phiCorr, phiCorrB, psiMax,
                        psiMin, nLimiterIter, rho, rho0, Sp, Su)
    return phiBD + lam * phiCorr, phiBDB + lamB * phiCorrB


def limit_corr(nCells, lower, upper, bFaceCells, V, rDeltaT, psi, psiB, phiB, phiCorr, phiCorrB, psiMax, psiMin, nLimiterIter=3,
               rho=None, Sp=None, Su=None, extremaCoeff=0.0):
    """MULES::limitCorr (CMULESTemplates.C:706-761): phiCorr *= lambda, lambda from limiterCorr"""
    lam, lamB = limiter(nCells, lower, upper, bFaceCells, V, rDeltaT, psi, psi, psiB, np.zeros(len(lower)), phiB, phiCorr, phiCorrB,
                        psiMax, psiMin, nLimiterIter, rho, None, Sp, Su, corr=True, extremaCoeff=extremaCoeff)
    return _a(phiCorr) * lam, _a(phiCorrB) * lamB


def correct(nCells, lower, upper, bFaceCells, V, rDeltaT, psi, phiCorr, phiCorrB, rho=None, Sp=None, Su=None):
    """MULES::correct (CMULESTemplates.C:35-75): psi = (rho*psi*rDeltaT + Su - surfaceIntegrate(phiCorr))/(rho*rDeltaT - Sp)"""
    sI = surface_integrate(nCells, lower, upper, bFaceCells, V, phiCorr, phiCorrB)
    num = (_a(psi) if rho is None else _a(rho) * _a(psi)) * rDeltaT
    if Su is not None:
        num = num + _a(Su)
    num = num - sI
    den = rDeltaT if rho is None else _a(rho) * rDeltaT
    if Sp is not None:
        den = den - _a(Sp)
    return num / den


def surface_integrate(nCells, lower, upper, bFaceCells, V, ssf, bssf):
    inc = Incidence(nCells, lower, upper, bFaceCells)
    allf = np.concatenate([_a(ssf), _a(bssf)])
    s = np.zeros(nCells)
    for k in range(inc.width):
        f, kind, ok = inc.face[:, k], inc.kind[:, k], inc.valid[:, k]
        s = np.where(ok, np.where(kind != 1, s + allf[f], s - allf[f]), s)
    return s / _a(V)


def explicit_solve(nCells, lower, upper, bFaceCells, V, rDeltaT, psi0, phiPsi, phiPsiB, rho=None, rho0=None, Sp=None, Su=None):
    sI = surface_integrate(nCells, lower, upper, bFaceCells, V, phiPsi, phiPsiB)
    r0 = rho0 if rho0 is not None else rho
    num = (_a(psi0) if r0 is None else _a(r0) * _a(psi0)) * rDeltaT
    if Su is not None:
        num = num + _a(Su)
    num = num - sI
    den = rDeltaT if rho is None else _a(rho) * rDeltaT
    if Sp is not None:
        den = den - _a(Sp)
    return num / den


# ---- the reference's own MULESTemplates.C (oracle/_ref/libref_mules.so) ----
def reference_available():
    from oracle import ref_ldu
    return os.path.exists(_LIB) and os.path.exists(ref_ldu._LIB_LDUADDR)


def reference(mode, nCells, lower, upper, patchStart, bFaceCells, V, rDeltaT, psi, psi0, psiB, a, b, psiMax=1.0, psiMin=0.0,
              nLimiterIter=3, rho=None, rho0=None, Sp=None, Su=None, nCoupledPatches=0, lambda0=None, extremaCoeff=0.0):
    """mode 0: MULES::limiter(a = phiBD, b = phiCorr) -> allLambda (lambda0: the starting limiter, default 1; the trailing
    nCoupledPatches patches answer coupled() and psiB holds their patchNeighbourField(); syncFaceList is a no-op in the harness); 1: MULES::limit(a = phi, b = phiPsi) -> phiPsi;
    2: MULES::explicitSolve(a = phiPsi) -> psi; 3: MULES::limiterCorr(a = phi, b = phiCorr) -> allLambda; 4: MULES::limitCorr -> the
    limited phiCorr; 5: MULES::correct(b = phiCorr) -> psi.  a, b: internal faces followed by the boundary faces in patch order.
    The patch sort addressing comes from the reference's own lduAddressing.C (ref_ldu.ldu_addressing)."""
    global _lib
    from oracle import ref_ldu
    if _lib is None:
        _lib = C.CDLL(_LIB)
    i32 = lambda x: np.ascontiguousarray(x, np.int32)
    f64 = lambda x: None if x is None else np.ascontiguousarray(x, np.float64)
    p = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
    l, u, ps, bfc = i32(lower), i32(upper), i32(patchStart), i32(bFaceCells)
    ad = ref_ldu.ldu_addressing(nCells, l, u, ps, bfc)
    nP = len(ps) - 1
    scs = i32(np.concatenate([[0], np.cumsum([len(x) for x in ad["patchSortCells"]])]))
    sc = i32(np.concatenate(ad["patchSortCells"])) if nP else i32([])
    sa = i32(np.concatenate(ad["patchSortAddr"])) if nP else i32([])
    ss = i32(np.concatenate(ad["patchSortStart"])) if nP else i32([])
    os_, ls, lo = i32(ad["ownerStart"]), i32(ad["losortStart"]), i32(ad["losort"])
    nF, nB = len(l), len(bfc)
    out = np.array(psi, np.float64).copy() if mode == 5 else np.zeros(nCells if mode == 2 else nF + nB)
    arrs = [f64(x) for x in (V, psi, psi0, psiB, rho, rho0, Sp, Su, a, b, lambda0)]
    Vv, psi_, psi0_, psiB_, rho_, rho0_, Sp_, Su_, a_, b_, l0_ = arrs
    _lib.ref_mules.argtypes = ([C.c_int] * 3 + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 7 + [C.c_double] + [C.c_void_p] * 9 +
                               [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_double])
    rc = _lib.ref_mules(mode, int(nCells), nF, p(l), p(u), p(os_), p(ls), p(lo), nP, p(ps), p(bfc), p(scs), p(sc), p(sa), p(ss),
                        p(Vv), float(rDeltaT), p(psi_), p(psi0_), p(psiB_), p(rho_), p(rho0_), p(Sp_), p(Su_), p(a_), p(b_),
                        float(psiMax), float(psiMin), int(nLimiterIter), p(out), int(nCoupledPatches), p(l0_), float(extremaCoeff))
    if rc != 0:
        raise RuntimeError("the reference code raised an error")
    return out


def reference_ranks(cases, patchStarts, nCoupledPatches, exchange, nLimiterIter):
    """MULES::limiter of a decomposed case through the reference's own code: one call with nLimiterIter = 1 per rank and sweep
    (the flux budgets of the first part do not depend on lambda, so k calls of one sweep = one call of k sweeps), the minimum with
    the other side's values in between -- what syncFaceList does.  cases: limiter_steps keyword dicts (nCoupled faces trailing)."""
    lam = [np.ones(len(c["lower"]) + len(c["bFaceCells"])) for c in cases]
    for _ in range(nLimiterIter):
        for r, c in enumerate(cases):
            lam[r] = reference(0, c["nCells"], c["lower"], c["upper"], patchStarts[r], c["bFaceCells"], c["V"], c["rDeltaT"], c["psi"],
                               c["psi0"], c["psiB"], np.concatenate([c["phiBD"], c["phiBDB"]]),
                               np.concatenate([c["phiCorr"], c["phiCorrB"]]), c["psiMax"], c["psiMin"], 1, c.get("rho"), c.get("rho0"),
                               c.get("Sp"), c.get("Su"), nCoupledPatches[r], lam[r])
        theirs = exchange([l[len(l) - c["nCoupled"]:] for l, c in zip(lam, cases)])
        for r, c in enumerate(cases):
            k = c["nCoupled"]
            lam[r][len(lam[r]) - k:] = np.minimum(lam[r][len(lam[r]) - k:], theirs[r])
    return [(l[: len(c["lower"])], l[len(c["lower"]):]) for l, c in zip(lam, cases)]
