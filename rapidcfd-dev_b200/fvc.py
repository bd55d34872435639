"""Explicit finite-volume operators that are compositions of C-ABI kernels (host side: sequencing only, every field stays
on the device).  `capi` is rapidcfd-dev_b200.capi (or, in CPU dry runs of the sequencing, the oracle-backed stand-in of
tests/oracle_backend.py); `ops` its FieldOps.  Reference: FV/ = src/finiteVolume/.

  snGrad_correction   correctedSnGrad<Type>::correction for a scalar field (FV/finiteVolume/snGradSchemes/correctedSnGrad/
                      correctedSnGrad.C:44-75): nonOrthCorrectionVectors & interpolate(grad(vf)); the same expression is
                      gaussLaplacianScheme::gammaSnGradCorr with SfGammaCorr in place of the correction vectors
                      (gaussLaplacianScheme.C:92-130)
  laplacian           fvc::laplacian(gamma, vf) = fvc::div(gamma*snGrad(vf)*magSf) (gaussLaplacianSchemes.C:95-112), snGrad
                      with or without the non-orthogonal correction
"""


def snGrad_correction(capi, addr, ops, corrVecs, Sf, w, vf, bSf, bvf, V):
    """internal faces: corrVecs & interpolate(grad(vf)), grad = Gauss linear with the boundary face values bvf"""
    g = capi.fv_grad_linear(addr, 1, Sf, w, vf, bSf, bvf, V)
    return ops.dot3(corrVecs, capi.fv_interpolate_linear(addr, 3, w, g))


def laplacian(capi, addr, ops, nComp, gammaMagSf, deltaCoeffs, vf, bFlux, V, correction=None):
    """gammaMagSf = gamma*magSf on the internal faces; bFlux = the flux on the boundary faces [nBFaces*nComp] (from the boundary
    conditions' snGrad); correction: the snGrad correction of the internal faces (scalar fields), added before the product"""
    sn = capi.fv_sngrad(addr, nComp, deltaCoeffs, vf)
    if correction is not None:
        sn = ops.add(sn, correction, nComp, nComp)
    return capi.fv_surface_integrate(addr, nComp, ops.mul(gammaMagSf, sn, 1, nComp), bFlux, V, True, -1)
