"""The standard k-epsilon model's transport step sequenced over the C ABI (host side: sequencing only, every field stays on the
device): the production term, the two scalar transport equations -- each assembled from the coefficient-fill kernels and solved
by PBiCG + DILU through the fvMatrix glue, i.e. through the same boundary as every other solve -- the bounding of k and epsilon,
and nut.  `capi` is rapidcfd-dev_b200.capi (or the oracle-backed stand-in of tests/oracle_backend.py).

  kEpsilon::correct      src/turbulenceModels/incompressible/RAS/kEpsilon/kEpsilon.C:227-276
  DkEff / DepsilonEff    kEpsilon.H:123-138
  bound                  FV/cfdTools/general/bound/bound.C:33-72 (applied unconditionally: where no value lies below the bound the
                         expression returns the field unchanged, which is what the reference's `if (minVsf < lowerBound)` skips)
Scope: single domain, fixedValue boundary values of k and epsilon on every patch; no wall functions (`boundaryManipulate`, the
`epsilon` / `nut` wall-function patch fields); div schemes upwind | linear."""

SMALL = 1e-15


class KEpsilon:
    def __init__(self, capi, case, bMagSf, bDeltaCoeffs, k, epsilon, kB, epsB, Cmu=0.09, C1=1.44, C2=1.92, sigmaEps=1.3, kMin=SMALL,
                 epsilonMin=SMALL):
        """case: the IcoFoam object that owns the mesh arrays on the device (addr, Sf, magSf, w, delta, V, bSf, Ub, nu)"""
        assert not case.nC, "k-epsilon step: single domain"
        self.capi, self.case, self.ops = capi, case, case.ops
        t = case._t
        self.bMagSf, self.bDelta = t(bMagSf), t(bDeltaCoeffs)
        self.k, self.epsilon, self.kB, self.epsB = t(k), t(epsilon), t(kB), t(epsB)
        self.Cmu, self.C1, self.C2, self.sigmaEps, self.kMin, self.epsilonMin = Cmu, C1, C2, sigmaEps, kMin, epsilonMin
        self.mat = capi.LduMatrix(case.addr)
        o = self.ops
        # surfaceSum(magSf): the denominator of fvc::average, constant
        self.sumMagSf = capi.fv_surface_integrate(case.addr, 1, case.magSf, self.bMagSf, case.V, False, 1)
        self.zeroB = o.smul(0.0, self.bMagSf)
        # constructor body (kEpsilon.C:134-141)
        self.k, self.kB = self.bound(self.k, self.kB, kMin)
        self.epsilon, self.epsB = self.bound(self.epsilon, self.epsB, epsilonMin)
        self.update_nut()

    def update_nut(self):
        o = self.ops
        self.nut = o.div(o.smul(self.Cmu, o.mul(self.k, self.k)), self.epsilon)
        self.nutB = o.div(o.smul(self.Cmu, o.mul(self.kB, self.kB)), self.epsB)

    def bound(self, vsf, vb, lowerBound):
        capi, c, o = self.capi, self.case, self.ops
        mx, mxb = o.smax(vsf, lowerBound), o.smax(vb, lowerBound)
        face = capi.fv_interpolate_linear(c.addr, 1, c.w, mx)
        num = capi.fv_surface_integrate(c.addr, 1, o.mul(c.magSf, face), o.mul(self.bMagSf, mxb), c.V, False, 1)
        av = o.div(num, self.sumMagSf)
        return o.smax(o.bmax(vsf, o.mul(av, o.pos(o.neg(vsf)))), lowerBound), mxb

    def _transport(self, psi, psiB, gamma, gammaB, su, sp, phi, bphi, rDeltaT, divScheme, alpha, ctl):
        capi, c, o = self.capi, self.case, self.ops
        a = c.addr
        ddtDiag, ddtSource = o.smul(rDeltaT, c.V), o.mul(o.smul(rDeltaT, psi), c.V)
        if divScheme not in ("upwind", "linear"):
            raise ValueError(f"Unknown discretisation scheme {divScheme}\n\nValid schemes are :\n(linear upwind)")
        wConv = capi.fv_limited_weights(c.ctx, phi) if divScheme == "upwind" else c.w
        cLower, cUpper, cDiag = capi.fv_convection_fill(a, wConv, phi)
        gf = capi.fv_interpolate_linear(a, 1, c.w, gamma)
        lUpper, lDiag = capi.fv_laplacian_fill(a, c.delta, o.mul(gf, c.magSf))
        diag, upper, lower = o.sub(o.add(ddtDiag, cDiag), lDiag), o.sub(cUpper, lUpper), o.sub(cLower, lUpper)
        gb = o.mul(gammaB, self.bMagSf)
        lIc, lBc = o.mul(gb, o.neg(self.bDelta)), o.mul(o.neg(gb), o.mul(self.bDelta, psiB))
        ic, bc = o.sub(self.zeroB, lIc), o.sub(o.mul(o.neg(bphi), psiB), lBc)
        diag = o.add(diag, o.mul(c.V, sp))
        source = o.add(ddtSource, o.mul(c.V, su))
        self.mat.set(diag, upper, lower)
        new = psi.clone()
        eqn = capi.FvMatrix(self.mat, 1, diag, source, new, c.V, ic, bc)
        if alpha is not None:
            eqn.relax(alpha)                     # in place on diag / source (fvMatrix.C:1088-1345) ...
            self.mat.set(diag, upper, lower)     # ... and the matrix holds copies: it takes the relaxed diagonal here
        perf = eqn.solve("PBiCG", "DILU", **ctl)[0]
        return new, perf

    def correct(self, U, phi, bphi, deltaT, divScheme="upwind", alphaEps=None, alphaK=None, controls=None):
        """kEpsilon::correct(): returns (perf of the epsilon solve, perf of the k solve)"""
        capi, c, o = self.capi, self.case, self.ops
        ctl = controls or dict(tolerance=1e-10, relTol=0.0)
        gradU = capi.fv_grad_linear(c.addr, 3, c.Sf, c.w, U, c.bSf, c.Ub, c.V)
        self.G = G = o.mul(o.smul(2.0, self.nut), o.symm_magsqr(gradU))
        rDeltaT = 1.0 / deltaT
        nu = c.nu
        self.epsilon, pe = self._transport(self.epsilon, self.epsB, o.sadd(o.sdiv(self.nut, self.sigmaEps), nu),
                                           o.sadd(o.sdiv(self.nutB, self.sigmaEps), nu),
                                           o.div(o.mul(o.smul(self.C1, G), self.epsilon), self.k),
                                           o.div(o.smul(self.C2, self.epsilon), self.k), phi, bphi, rDeltaT, divScheme, alphaEps, ctl)
        self.epsilon, self.epsB = self.bound(self.epsilon, self.epsB, self.epsilonMin)
        self.k, pk = self._transport(self.k, self.kB, o.sadd(self.nut, nu), o.sadd(self.nutB, nu), G, o.div(self.epsilon, self.k), phi,
                                     bphi, rDeltaT, divScheme, alphaK, ctl)
        self.k, self.kB = self.bound(self.k, self.kB, self.kMin)
        self.update_nut()
        return pe, pk

    def close(self):
        self.mat.close()
