"""Explicit MULES as compositions of C-ABI kernels (host side: sequencing only, every field stays on the device).  `capi` is
rapidcfd-dev_b200.capi (or, in CPU dry runs of the sequencing, the oracle-backed stand-in of tests/oracle_backend.py); `ops`
its FieldOps.  Reference: FV/fvMatrices/solvers/MULES/MULESTemplates.C.

  limit            MULES::limit (:748-813): phiBD = upwind<scalar>(mesh, phi).flux(psi); phiCorr = phiPsi - phiBD; lambda = 1;
                   MULES::limiter (b200ldu_mules_limiter); phiPsi = phiBD + lambda*phiCorr
  explicit_solve   MULES::explicitSolve (:36-78): psi = (rho0*psi0*rDeltaT + Su - surfaceIntegrate(phiPsi))/(rho*rDeltaT - Sp)
  limit_corr       MULES::limitCorr (CMULESTemplates.C:706-761): phiCorr *= lambda, lambda from MULES::limiterCorr
                   (b200ldu_mules_limiter_corr)
  correct          MULES::correct (CMULESTemplates.C:35-75): psi = (rho*psi*rDeltaT + Su - surfaceIntegrate(phiCorr))/(rho*rDeltaT - Sp)
rho / rho0 None: geometricOneField; Sp / Su None: zeroField -- the operations the reference's one / zero algebra drops are not
issued.  Static mesh; boundary faces = the faces of fv_boundary_set in patch order, coupled (processor / cyclic) patch faces last."""


def _cat(a, b):
    import torch
    return torch.cat([a, b])


def upwind_flux(capi, addr, ops, phi, phiB, psi, psiB, nCoupled=0, cfc=None):
    """upwind<scalar>::flux (upwind.H:86-103 weights pos(faceFlux); surfaceInterpolationScheme.C:176-184 faceFlux*interpolate).
    Coupled patch faces (the trailing nCoupled boundary faces, psiB = patchNeighbourField, cfc = their face cells):
    w*patchInternalField + (1 - w)*patchNeighbourField (surfaceInterpolationScheme.C:300-312)"""
    w = capi.fv_limited_weights(addr.ctx, phi)
    bd = ops.mul(phi, capi.fv_interpolate_linear(addr, 1, w, psi))
    if not nCoupled:
        return bd, ops.mul(phiB, psiB)
    nW = phiB.numel() - nCoupled
    wc = capi.fv_limited_weights(addr.ctx, phiB[nW:])
    sf = ops.add(ops.mul(wc, ops.gather(cfc, psi)), ops.mul(ops.rsub(1.0, wc), psiB[nW:]))
    return bd, _cat(ops.mul(phiB[:nW], psiB[:nW]), ops.mul(phiB[nW:], sf))


def limit(capi, addr, ops, V, rDeltaT, psi, psi0, psiB, phi, phiB, phiPsi, phiPsiB, psiMax, psiMin, nLimiterIter=3,
          rho=None, rho0=None, Sp=None, Su=None, nCoupled=0, cfc=None):
    """returns the limited (phiPsi, phiPsiB).  Decomposed meshes: the boundary faces of fv_boundary_set end with the nCoupled coupled
    patch faces (face cells cfc), psiB there = fv_patch_neighbour_field(psi); collective over the ranks"""
    phiBD, phiBDB = upwind_flux(capi, addr, ops, phi, phiB, psi, psiB, nCoupled, cfc)
    phiCorr, phiCorrB = ops.sub(phiPsi, phiBD), ops.sub(phiPsiB, phiBDB)
    lam, lamB = capi.mules_limiter(addr, V, rDeltaT, psi, psi0, psiB, phiBD, phiBDB, phiCorr, phiCorrB, psiMax, psiMin,
                                   nLimiterIter, rho, rho0 if rho0 is not None else rho, Sp, Su, nCoupled)
    return ops.add(phiBD, ops.mul(lam, phiCorr)), ops.add(phiBDB, ops.mul(lamB, phiCorrB))


def explicit_solve(capi, addr, ops, V, rDeltaT, psi0, phiPsi, phiPsiB, rho=None, rho0=None, Sp=None, Su=None):
    """returns the new psi"""
    sI = capi.fv_surface_integrate(addr, 1, phiPsi, phiPsiB, V, True, -1)
    r0 = rho0 if rho0 is not None else rho
    num = ops.smul(rDeltaT, psi0 if r0 is None else ops.mul(r0, psi0))
    if Su is not None:
        num = ops.add(num, Su)
    num = ops.sub(num, sI)
    if rho is None and Sp is None:
        return ops.sdiv(num, rDeltaT)
    den = ops.smul(rDeltaT, rho) if rho is not None else None
    if Sp is not None:
        den = ops.sub(den, Sp) if den is not None else ops.rsub(rDeltaT, Sp)
    return ops.div(num, den)


def limit_corr(capi, addr, ops, V, rDeltaT, psi, psiB, phiB, phiCorr, phiCorrB, psiMax, psiMin, nLimiterIter=3, rho=None, Sp=None,
               Su=None, extremaCoeff=0.0, nCoupled=0):
    """returns the limited correction (lambda*phiCorr, lambdaB*phiCorrB); phiB: boundary values of the total flux"""
    lam, lamB = capi.mules_limiter_corr(addr, V, rDeltaT, psi, psiB, phiB, phiCorr, phiCorrB, psiMax, psiMin, nLimiterIter, rho, Sp, Su,
                                        extremaCoeff, nCoupled)
    return ops.mul(phiCorr, lam), ops.mul(phiCorrB, lamB)


def correct(capi, addr, ops, V, rDeltaT, psi, phiCorr, phiCorrB, rho=None, Sp=None, Su=None):
    """returns the corrected psi"""
    sI = capi.fv_surface_integrate(addr, 1, phiCorr, phiCorrB, V, True, -1)
    num = ops.smul(rDeltaT, psi if rho is None else ops.mul(rho, psi))
    if Su is not None:
        num = ops.add(num, Su)
    num = ops.sub(num, sI)
    if rho is None and Sp is None:
        return ops.sdiv(num, rDeltaT)
    den = ops.smul(rDeltaT, rho) if rho is not None else None
    if Sp is not None:
        den = ops.sub(den, Sp) if den is not None else ops.rsub(rDeltaT, Sp)
    return ops.div(num, den)
