"""Explicit MULES as compositions of C-ABI kernels (host side: sequencing only, every field stays on the device).  `capi` is
rapidcfd-dev_b200.capi (or, in CPU dry runs of the sequencing, the oracle-backed stand-in of tests/oracle_backend.py); `ops`
its FieldOps.  Reference: FV/fvMatrices/solvers/MULES/MULESTemplates.C.

  limit            MULES::limit (:748-813): phiBD = upwind<scalar>(mesh, phi).flux(psi); phiCorr = phiPsi - phiBD; lambda = 1;
                   MULES::limiter (b200ldu_mules_limiter); phiPsi = phiBD + lambda*phiCorr
  explicit_solve   MULES::explicitSolve (:36-78): psi = (rho0*psi0*rDeltaT + Su - surfaceIntegrate(phiPsi))/(rho*rDeltaT - Sp)
rho / rho0 None: geometricOneField; Sp / Su None: zeroField -- the operations the reference's one / zero algebra drops are not
issued.  Static mesh; boundary faces = the non-coupled faces of fv_boundary_set, in patch order."""


def upwind_flux(capi, addr, ops, phi, phiB, psi, psiB):
    """upwind<scalar>::flux (upwind.H:86-103 weights pos(faceFlux); surfaceInterpolationScheme.C:176-184 faceFlux*interpolate)"""
    w = capi.fv_limited_weights(addr.ctx, phi)
    return ops.mul(phi, capi.fv_interpolate_linear(addr, 1, w, psi)), ops.mul(phiB, psiB)


def limit(capi, addr, ops, V, rDeltaT, psi, psi0, psiB, phi, phiB, phiPsi, phiPsiB, psiMax, psiMin, nLimiterIter=3,
          rho=None, rho0=None, Sp=None, Su=None):
    """returns the limited (phiPsi, phiPsiB)"""
    phiBD, phiBDB = upwind_flux(capi, addr, ops, phi, phiB, psi, psiB)
    phiCorr, phiCorrB = ops.sub(phiPsi, phiBD), ops.sub(phiPsiB, phiBDB)
    lam, lamB = capi.mules_limiter(addr, V, rDeltaT, psi, psi0, psiB, phiBD, phiBDB, phiCorr, phiCorrB, psiMax, psiMin,
                                   nLimiterIter, rho, rho0 if rho0 is not None else rho, Sp, Su)
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
