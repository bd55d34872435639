"""CPU restatement of one icoFoam time step -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/applications/solvers/incompressible/icoFoam/icoFoam.C:55-103 statement by statement, each
field expression evaluated as the reference's gpuField operators evaluate it: one rounded elementwise operation per
operator, in the written order (numpy does the same).  Schemes: ddt Euler, div Gauss linear, laplacian Gauss linear
uncorrected (orthogonal meshes), grad Gauss linear, interpolate linear -- the cavity tutorial's fvSchemes.  Boundary
conditions: U fixedValue on every boundary face, p zeroGradient (the lid-driven cavity); single domain.  PARITY
UNPINNED (the application needs the whole library); checked by the physics it has to produce
(tests/test_oracle_piso.py).  FV/ = src/finiteVolume/.
"""
import numpy as np

from . import fvm_oracle as fo

SMALL = 1e-15   # src/OpenFOAM/primitives/Scalar/doubleScalar/doubleScalar.H


class Cavity:
    """mesh + fields of the case; arrays in OpenFOAM order (faces: owner-sorted upper triangle; boundary faces: all
    patches concatenated in patch order)."""

    def __init__(self, orc, nCells, lower, upper, Sf, magSf, weights, deltaCoeffs, V, bFaceCells, bSf, bMagSf,
                 bDeltaCoeffs, Ub, nu, deltaT, pRefCell=0, pRefValue=0.0):
        self.orc = orc
        self.addr = orc.Addr(nCells, lower, upper)
        self.n, self.lower, self.upper = int(nCells), np.asarray(lower), np.asarray(upper)
        self.Sf, self.magSf, self.w, self.delta, self.V = (np.asarray(x, float) for x in (Sf, magSf, weights, deltaCoeffs, V))
        self.bfc = np.asarray(bFaceCells, np.int64)
        self.bSf, self.bMagSf, self.bDelta, self.Ub = (np.asarray(x, float) for x in (bSf, bMagSf, bDeltaCoeffs, Ub))
        self.nu, self.deltaT, self.pRefCell, self.pRefValue = float(nu), float(deltaT), int(pRefCell), float(pRefValue)
        self.U = np.zeros((self.n, 3))
        self.p = np.zeros(self.n)
        # createPhi.H: phi = linearInterpolate(U) & mesh.Sf()
        self.phi = self.flux_of(self.U)
        self.bphi = self._dot(self.Ub, self.bSf)

    # ---- field expressions ----------------------------------------------------------------------------
    @staticmethod
    def _dot(a, b):            # Vector & Vector: x*x + y*y + z*z, left to right (VectorI.H)
        return (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]

    def interpolate(self, vf):
        """linear: w*vf[own] + (1 - w)*vf[nei] (surfaceInterpolationScheme.C:159-240)"""
        w = self.w if vf.ndim == 1 else self.w[:, None]
        return w * vf[self.lower] + (1 - w) * vf[self.upper]

    def flux_of(self, U):
        return self._dot(self.interpolate(U), self.Sf)

    def grad(self, p, pb):
        """fvc::grad = gaussGrad(linear): gradf(interpolate(p)) (gaussGrad.C:143-271); the boundary correction of
        correctBoundaryConditions only touches the patch values of the gradient, which nothing here reads"""
        g = self.orc.gauss_grad(self.addr, self.Sf.ravel(), self.interpolate(p), self.bfc, self.bSf.ravel(), pb, self.V, 1)
        return np.asarray(g).reshape(self.n, 3)

    def div(self, phi, bphi):
        """fvc::div(flux) = surfaceIntegrate (fvcDiv.C, fvcSurfaceIntegrate.C:138-203)"""
        return np.asarray(self.orc.surface_integrate(self.addr, phi, self.bfc, bphi, self.V, 1))

    # ---- one time step: icoFoam.C:55-103 --------------------------------------------------------------
    def step(self, nCorr=2, nNonOrthCorr=0, UControls=None, pControls=None, momentumPredictor=True,
             USolver=("PBiCG", "DILU"), pSolver=("PCG", "DIC"), gamg=None):
        """USolver / pSolver: (solver, preconditioner or smoother) as fvSolution names them; gamg: the cached
        agglomeration (orc.Gamg) when pSolver is GAMG"""
        orc = self.orc
        U0, phi0 = self.U.copy(), self.phi.copy()                    # oldTime fields
        rDeltaT = 1.0 / self.deltaT
        # fvm::ddt(U): EulerDdtScheme.C:331-361
        ddtDiag = rDeltaT * self.V
        ddtSource = (rDeltaT * U0) * self.V[:, None]
        # fvm::div(phi, U): gaussConvectionScheme.C:76-115 (lower = -w*phi, upper = lower + phi, negSumDiag; fixedValue
        # patches: internalCoeffs = 0, boundaryCoeffs = -patchFlux*U_b)
        cLower, cUpper, cDiag = orc.convection_fill(self.addr, self.w, self.phi)
        cIc = self.bphi[:, None] * np.zeros((len(self.bfc), 3))
        cBc = (-self.bphi)[:, None] * self.Ub
        # fvm::laplacian(nu, U): gaussLaplacianSchemes.C:39-93, gaussLaplacianScheme.C:46-89
        gammaMagSf = self.nu * self.magSf
        lUpper, lDiag = orc.laplacian_fill(self.addr, self.delta, gammaMagSf)
        lIc, lBc = fo.fixedValue_laplacian_coeffs(self.nu * self.bMagSf, self.bDelta, self.Ub)
        # UEqn = ddt + div - laplacian (fvMatrix.C operator+ / operator-: coefficient arrays added then subtracted)
        diag = (ddtDiag + cDiag) - lDiag
        upper = cUpper - lUpper
        lower = cLower - lUpper
        source = ddtSource
        ic = cIc - lIc
        bc = cBc - lBc
        perfs = {}
        pb = self.p[self.bfc]                                         # zeroGradient
        if momentumPredictor:
            # solve(UEqn == -fvc::grad(p)): source += V*(-grad p) (fvMatrix.C operator==, operator-)
            gradP = self.grad(self.p, pb)
            src = source + self.V[:, None] * (-gradP)
            UEqn = fo.FvMatrix(orc, self.addr, 3, diag, upper, lower, src, self.U, self.V, self.bfc, ic, bc)
            self.U, perfs["U"], _ = UEqn.solve(USolver[0], USolver[1], **(UControls or dict(tolerance=1e-5, relTol=0.0)))
        cont = []
        for corr in range(nCorr):
            UEqn = fo.FvMatrix(orc, self.addr, 3, diag, upper, lower, source, self.U, self.V, self.bfc, ic, bc)
            rAU = 1.0 / UEqn.A()
            HbyA = rAU[:, None] * UEqn.H()
            # phiHbyA = (interpolate(HbyA) & Sf) + interpolate(rAU)*ddtCorr(U, phi)
            phiCorr = phi0 - self._dot(self.Sf, self.interpolate(U0))             # EulerDdtScheme.C:536-539
            coeff = 1.0 - np.minimum(np.abs(phiCorr) / (np.abs(phi0) + SMALL), 1.0)   # ddtScheme.C:139-152
            ddtCorr = (coeff * rDeltaT) * phiCorr
            phiHbyA = self.flux_of(HbyA) + self.interpolate(rAU) * ddtCorr
            bphiHbyA = self._dot(self.Ub, self.bSf)        # fixedValue: HbyA_b = U_b, coupling coefficient 0 (:156-162)
            # adjustPhi: closed domain, the boundary flux is U_b & Sf_b = 0: nothing to adjust (adjustPhi.C)
            for nonOrth in range(nNonOrthCorr + 1):
                # pEqn: fvm::laplacian(rAU, p) == fvc::div(phiHbyA)
                rAUf = self.interpolate(rAU)
                pUpper, pDiag = orc.laplacian_fill(self.addr, self.delta, rAUf * self.magSf)
                pSource = np.zeros(self.n) + self.V * self.div(phiHbyA, bphiHbyA)
                zero = np.zeros((len(self.bfc), 1))
                pEqn = fo.FvMatrix(orc, self.addr, 1, pDiag, pUpper, None, pSource, self.p, self.V, self.bfc, zero, zero)
                pEqn.setReference(self.pRefCell, self.pRefValue)
                psi, perf, _ = pEqn.solve(pSolver[0], pSolver[1], gamg, **(pControls or dict(tolerance=1e-6, relTol=0.0)))
                self.p = psi[:, 0]
                perfs.setdefault("p", []).append(perf[0])
                if nonOrth == nNonOrthCorr:
                    pEqn.psi = psi
                    internal, boundary, _ = pEqn.flux()
                    self.phi = phiHbyA - internal[:, 0]
                    self.bphi = bphiHbyA - boundary[:, 0]
            # continuityErrs.H
            contErr = self.div(self.phi, self.bphi)
            cont.append((self.deltaT * (np.abs(contErr) * self.V).sum() / self.V.sum(),
                         self.deltaT * (contErr * self.V).sum() / self.V.sum()))
            # U = HbyA - rAU*fvc::grad(p); U.correctBoundaryConditions() (fixedValue: unchanged)
            self.U = HbyA - rAU[:, None] * self.grad(self.p, self.p[self.bfc])
        return perfs, cont


def cavity_from_hex(orc, meshmod, n, nu=0.01, deltaT=None, lid=(1.0, 0.0, 0.0)):
    """the lid-driven cavity on the synthetic n^3 hex mesh of rapidcfd-dev_b200/mesh.py (patch `movingWall` = +y)"""
    m = meshmod.hex_mesh(n)
    bfc = np.concatenate([p.faceCells for p in m.patches]).astype(np.int32)
    bSf = np.concatenate([p.Sf for p in m.patches])
    Ub = np.concatenate([np.tile(lid if p.name == "movingWall" else (0.0, 0.0, 0.0), (len(p.faceCells), 1))
                         for p in m.patches])
    nB = len(bfc)
    deltaT = deltaT if deltaT is not None else 0.5 * m.h / max(abs(v) for v in lid)     # Co = 0.5 at the lid
    return m, Cavity(orc, m.nCells, m.lower, m.upper, m.Sf(), m.magSf(), m.weights(), m.deltaCoeffs(), m.volumes(), bfc,
                     bSf, np.full(nB, m.h * m.h), np.full(nB, 2.0 / m.h), Ub, nu, deltaT)
