"""One icoFoam time step on the device: the statement sequence of
/root/reference/applications/solvers/incompressible/icoFoam/icoFoam.C:55-103 over the C ABI -- the face-sum and
coefficient kernels of csrc/fv.cu, the fvMatrix glue of csrc/fvmatrix.cu, the linear solvers, and between them the
gpuField operator set of csrc/fieldops.cu composed in the reference's order (one rounded operation per operator, as
the reference's expression templates evaluate).  Host side: this module only sequences calls; every field stays on
the device.  Schemes and boundary conditions of the cavity tutorial: ddt Euler, div / laplacian / grad Gauss linear
(uncorrected: orthogonal mesh), U fixedValue on every boundary face, p zeroGradient; single domain.

The fixedValue / zeroGradient coefficient arrays are what the reference's boundary-condition objects hand to the
schemes (fixedValueFvPatchField.C:113-146, zeroGradientFvPatchField.C); they are constant over the run and built once
on the host at construction.
"""
import importlib

import numpy as np

SMALL = 1e-15   # doubleScalar.H


class IcoFoam:
    def __init__(self, capi, ctx, torch, nCells, lower, upper, Sf, magSf, weights, deltaCoeffs, V, bFaceCells, bSf,
                 bMagSf, bDeltaCoeffs, Ub, nu, deltaT, pRefCell=0, pRefValue=0.0, cellCentres=None, addr=None):
        self.capi, self.ctx, self.torch = capi, ctx, torch
        self.n, self.nF = int(nCells), len(lower)
        self.addr = addr if addr is not None else capi.LduAddressing(ctx, nCells, lower, upper, cellCentres=cellCentres)
        bfc = np.ascontiguousarray(bFaceCells, dtype=np.int32)
        self.nB = len(bfc)
        capi.fv_boundary_set(self.addr, bfc)
        self.ops = capi.FieldOps(ctx)
        t = self._t
        self.Sf, self.magSf, self.w, self.delta, self.V = t(Sf), t(magSf), t(weights), t(deltaCoeffs), t(V)
        self.bfc = torch.from_numpy(np.array(bfc)).to(ctx.device)
        self.bSf, self.Ub = t(bSf), t(Ub)
        self.nu, self.deltaT, self.pRefCell, self.pRefValue = float(nu), float(deltaT), int(pRefCell), float(pRefValue)
        # boundary-condition coefficients (constant): fvm::laplacian(nu, U) on fixedValue patches
        # (gaussLaplacianScheme.C:66-86, fixedValueFvPatchField.C:136-146)
        g = (nu * np.asarray(bMagSf, float))[:, None]
        d = np.asarray(bDeltaCoeffs, float)[:, None]
        Ubh = np.asarray(Ub, float).reshape(self.nB, 3)
        self.lIc, self.lBc = t(g * (-d) * np.ones_like(Ubh)), t((-g) * (d * Ubh))
        self.zeroB3, self.zeroB1 = t(np.zeros((self.nB, 3))), t(np.zeros(self.nB))
        self.U, self.p = t(np.zeros((self.n, 3))), t(np.zeros(self.n))
        # createPhi.H: phi = linearInterpolate(U) & mesh.Sf()
        self.phi = capi.fv_flux_linear(self.addr, self.Sf, self.w, self.U)
        self.bphi = self.ops.dot3(self.Ub, self.bSf)
        self.bUSf = self.ops.dot3(self.Ub, self.bSf)
        self.matU, self.matP = capi.LduMatrix(self.addr), capi.LduMatrix(self.addr)

    def _t(self, a):
        return self.torch.from_numpy(np.array(a, dtype=np.float64).ravel()).to(self.ctx.device)

    def grad(self, p):
        """fvc::grad(p), Gauss linear, p zeroGradient at the boundary"""
        pb = self.ops.gather(self.bfc, p, 1)
        return self.capi.fv_grad_linear(self.addr, 1, self.Sf, self.w, p, self.bSf, pb, self.V)

    def div(self, phi, bphi):
        return self.capi.fv_surface_integrate(self.addr, 1, phi, bphi, self.V, True, -1)

    def step(self, nCorr=2, nNonOrthCorr=0, UControls=None, pControls=None, momentumPredictor=True):
        """icoFoam.C:55-103; returns ({"U": [Perf x3], "p": [Perf per pressure solve]}, continuity errors)"""
        capi, o, a = self.capi, self.ops, self.addr
        U0, phi0 = self.U.clone(), self.phi.clone()
        rDeltaT = 1.0 / self.deltaT
        # fvm::ddt(U)
        ddtDiag = o.smul(rDeltaT, self.V)
        ddtSource = o.mul(o.smul(rDeltaT, U0), self.V, 3, 1)
        # fvm::div(phi, U)
        cLower, cUpper, cDiag = capi.fv_convection_fill(a, self.w, self.phi)
        cBc = o.mul(o.neg(self.bphi), self.Ub, 1, 3)
        # fvm::laplacian(nu, U)
        lUpper, lDiag = capi.fv_laplacian_fill(a, self.delta, o.smul(self.nu, self.magSf))
        # UEqn = ddt + div - laplacian
        diag = o.sub(o.add(ddtDiag, cDiag), lDiag)
        upper = o.sub(cUpper, lUpper)
        lower = o.sub(cLower, lUpper)
        source = ddtSource
        ic = o.sub(self.zeroB3, self.lIc)
        bc = o.sub(cBc, self.lBc)
        self.matU.set(diag, upper, lower)
        perfs = {}
        if momentumPredictor:
            # solve(UEqn == -fvc::grad(p))
            src = o.add(source, o.mul(self.V, o.neg(self.grad(self.p)), 1, 3))
            UEqn = capi.FvMatrix(self.matU, 3, diag, src, self.U, self.V, ic, bc)
            perfs["U"] = UEqn.solve("PBiCG", "DILU", **(UControls or dict(tolerance=1e-5, relTol=0.0)))
        cont = []
        for corr in range(nCorr):
            UEqn = capi.FvMatrix(self.matU, 3, diag, source, self.U, self.V, ic, bc)
            rAU = o.rdiv(1.0, UEqn.A())
            HbyA = o.mul(rAU, UEqn.H(), 1, 3)
            # phiHbyA = (interpolate(HbyA) & Sf) + interpolate(rAU)*ddtCorr(U, phi)
            phiCorr = o.sub(phi0, capi.fv_flux_linear(a, self.Sf, self.w, U0))
            coeff = o.rsub(1.0, o.smin(o.div(o.mag(phiCorr), o.sadd(o.mag(phi0), SMALL)), 1.0))
            ddtCorr = o.mul(o.smul(rDeltaT, coeff), phiCorr)
            rAUf = capi.fv_interpolate_linear(a, 1, self.w, rAU)
            phiHbyA = o.add(capi.fv_flux_linear(a, self.Sf, self.w, HbyA), o.mul(rAUf, ddtCorr))
            bphiHbyA = self.bUSf
            for nonOrth in range(nNonOrthCorr + 1):
                # pEqn: fvm::laplacian(rAU, p) == fvc::div(phiHbyA)
                pUpper, pDiag = capi.fv_laplacian_fill(a, self.delta, o.mul(rAUf, self.magSf))
                pSource = o.mul(self.V, self.div(phiHbyA, bphiHbyA))
                pEqn = capi.FvMatrix(self.matP, 1, pDiag, pSource, self.p, self.V, self.zeroB1, self.zeroB1)
                self.matP.set(pDiag, pUpper)
                pEqn.setReference(self.pRefCell, self.pRefValue)
                perfs.setdefault("p", []).extend(pEqn.solve("PCG", "DIC", **(pControls or dict(tolerance=1e-6, relTol=0.0))))
                if nonOrth == nNonOrthCorr:
                    internal, boundary, _ = pEqn.flux(self.nB)
                    self.phi = o.sub(phiHbyA, internal)
                    self.bphi = o.sub(bphiHbyA, boundary)
            contErr = self.div(self.phi, self.bphi)
            vol = float(self.V.sum())
            cont.append((self.deltaT * float((contErr.abs() * self.V).sum()) / vol,
                         self.deltaT * float((contErr * self.V).sum()) / vol))
            # U = HbyA - rAU*fvc::grad(p)
            self.U = o.sub(HbyA, o.mul(rAU, self.grad(self.p), 1, 3), 3, 3)
        return perfs, cont

    def close(self):
        self.matU.close()
        self.matP.close()
        self.addr.close()


def cavity(capi, ctx, torch, n, nu=0.01, deltaT=None, lid=(1.0, 0.0, 0.0)):
    """the lid-driven cavity on the synthetic n^3 hex mesh (mesh.py; patch `movingWall` = +y)"""
    meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
    m = meshmod.hex_mesh(n)
    bfc = np.concatenate([p.faceCells for p in m.patches]).astype(np.int32)
    bSf = np.concatenate([p.Sf for p in m.patches])
    Ub = np.concatenate([np.tile(lid if p.name == "movingWall" else (0.0, 0.0, 0.0), (len(p.faceCells), 1))
                         for p in m.patches])
    nB = len(bfc)
    deltaT = deltaT if deltaT is not None else 0.5 * m.h / max(abs(v) for v in lid)
    return m, IcoFoam(capi, ctx, torch, m.nCells, m.lower, m.upper, m.Sf(), m.magSf(), m.weights(), m.deltaCoeffs(),
                      m.volumes(), bfc, bSf, np.full(nB, m.h * m.h), np.full(nB, 2.0 / m.h), Ub, nu, deltaT,
                      cellCentres=m.cell_centres())
