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
    """One rank of the case.  Boundary faces: the fixedValue patches (`bFaceCells`, patch order) and, on a decomposed
    case, the processor patches (`cou*` arrays, flat in patch order, as given to the addressing).  The face-sum kernels
    see all of them as one boundary list; the fvMatrix glue gets zero internal / boundary coefficients for the
    processor faces there and their real coefficients through the matrix' interface coefficients."""

    def __init__(self, capi, ctx, torch, nCells, lower, upper, Sf, magSf, weights, deltaCoeffs, V, bFaceCells, bSf,
                 bMagSf, bDeltaCoeffs, Ub, nu, deltaT, pRefCell=0, pRefValue=0.0, cellCentres=None, addr=None,
                 couPatchStart=None, couFaceCells=None, neighbRank=None, couSf=None, couMagSf=None, couWeights=None,
                 couDeltaCoeffs=None, allsum=None):
        self.capi, self.ctx, self.torch, self.allsum = capi, ctx, torch, allsum
        self.n, self.nF = int(nCells), len(lower)
        cfc = np.zeros(0, np.int32) if couFaceCells is None else np.ascontiguousarray(couFaceCells, dtype=np.int32)
        self.nC = len(cfc)
        if addr is not None:
            self.addr = addr
        elif self.nC:
            self.addr = capi.LduAddressing(ctx, nCells, lower, upper, couPatchStart, cfc, neighbRank, cellCentres)
        else:
            self.addr = capi.LduAddressing(ctx, nCells, lower, upper, cellCentres=cellCentres)
        bfc = np.ascontiguousarray(bFaceCells, dtype=np.int32)
        self.nB = len(bfc)
        capi.fv_boundary_set(self.addr, np.concatenate([bfc, cfc]))
        self.ops = capi.FieldOps(ctx)
        t = self._t
        self.Sf, self.magSf, self.w, self.delta, self.V = t(Sf), t(magSf), t(weights), t(deltaCoeffs), t(V)
        self.bfc = torch.from_numpy(np.array(bfc)).to(ctx.device)
        self.cfc = torch.from_numpy(np.array(cfc)).to(ctx.device)
        self.bSf, self.Ub = t(bSf), t(Ub)
        self.cSf = t(np.zeros((0, 3)) if couSf is None else couSf)
        self.allSf = torch.cat([self.bSf, self.cSf])
        z = np.zeros(self.nC)
        self.cMagSf, self.cw, self.cDelta = (t(z if x is None else x) for x in (couMagSf, couWeights, couDeltaCoeffs))
        self.nu, self.deltaT, self.pRefCell, self.pRefValue = float(nu), float(deltaT), int(pRefCell), float(pRefValue)
        # boundary-condition coefficients (constant): fvm::laplacian(nu, U) on fixedValue patches
        # (gaussLaplacianScheme.C:66-86, fixedValueFvPatchField.C:136-146)
        g = (nu * np.asarray(bMagSf, float))[:, None]
        d = np.asarray(bDeltaCoeffs, float)[:, None]
        Ubh = np.asarray(Ub, float).reshape(self.nB, 3)
        self.lIc, self.lBc = t(g * (-d) * np.ones_like(Ubh)), t((-g) * (d * Ubh))
        self.zeroB3, self.zeroB1 = t(np.zeros((self.nB, 3))), t(np.zeros(self.nB + self.nC))
        self.zeroC3 = t(np.zeros((self.nC, 3)))
        self.U, self.p = t(np.zeros((self.n, 3))), t(np.zeros(self.n))
        # createPhi.H: phi = linearInterpolate(U) & mesh.Sf()
        self.phi = capi.fv_flux_linear(self.addr, self.Sf, self.w, self.U)
        self.bUSf = self.ops.dot3(self.Ub, self.bSf)
        self.bphi = self.bUSf.clone()
        self.cphi = self.ops.dot3(self.interpolate_coupled(self.U, 3), self.cSf) if self.nC else t(z)
        self.matU, self.matP = capi.LduMatrix(self.addr), capi.LduMatrix(self.addr)

    def _t(self, a):
        return self.torch.from_numpy(np.array(a, dtype=np.float64).ravel()).to(self.ctx.device)

    def interpolate_coupled(self, vf, nc):
        """processor faces: w*patchInternalField + (1 - w)*patchNeighbourField (surfaceInterpolationScheme.C:357-370)"""
        o = self.ops
        pnf = self.capi.fv_patch_neighbour_field(self.addr, nc, vf)
        return o.add(o.mul(self.cw, o.gather(self.cfc, vf, nc), 1, nc), o.mul(o.rsub(1.0, self.cw), pnf, 1, nc), nc, nc)

    def grad(self, p):
        """fvc::grad(p), Gauss linear, p zeroGradient at the walls, interpolated on the processor faces"""
        pb = self.ops.gather(self.bfc, p, 1)
        if self.nC:
            pb = self.torch.cat([pb, self.interpolate_coupled(p, 1)])
        return self.capi.fv_grad_linear(self.addr, 1, self.Sf, self.w, p, self.allSf, pb, self.V)

    def div(self, phi, bphi, cphi=None):
        cphi = self.cphi if cphi is None else cphi
        return self.capi.fv_surface_integrate(self.addr, 1, phi, self.torch.cat([bphi, cphi]) if self.nC else bphi,
                                              self.V, True, -1)

    def _gsum3(self, a, b, c):
        v = [float(a), float(b), float(c)]
        return v if self.allsum is None else [float(x) for x in self.allsum(np.array(v))]

    def step(self, nCorr=2, nNonOrthCorr=0, UControls=None, pControls=None, momentumPredictor=True,
             USolver=("PBiCG", "DILU"), pSolver=("PCG", "DIC"), gamg=None, divScheme="linear"):
        """divScheme: the interpolation scheme of `div(phi,U) Gauss <scheme>` in system/fvSchemes: "linear" or "upwind" (weights
        pos(phi) on the internal and the processor faces, upwind.H:120-123).
        icoFoam.C:55-103; returns ({"U": [Perf x3], "p": [Perf per pressure solve]}, continuity errors).
        USolver / pSolver: (solver, preconditioner or smoother) as system/fvSolution names them; gamg: the cached
        agglomeration (capi.GamgAgglomeration over self.addr) when pSolver is GAMG."""
        capi, o, a, cat, nC = self.capi, self.ops, self.addr, self.torch.cat, self.nC
        U0, phi0, cphi0 = self.U.clone(), self.phi.clone(), self.cphi.clone()
        rDeltaT = 1.0 / self.deltaT
        # fvm::ddt(U)
        ddtDiag = o.smul(rDeltaT, self.V)
        ddtSource = o.mul(o.smul(rDeltaT, U0), self.V, 3, 1)
        # fvm::div(phi, U)
        if divScheme not in ("linear", "upwind"):
            raise ValueError(f"Unknown discretisation scheme {divScheme}\n\nValid schemes are :\n(linear upwind)")
        upwind = divScheme == "upwind"
        wConv = capi.fv_limited_weights(self.ctx, self.phi) if upwind else self.w
        cLower, cUpper, cDiag = capi.fv_convection_fill(a, wConv, self.phi)
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
        ci = cb = None
        if nC:
            # processor patches (coupledFvPatchField.C:162-209): convection patchFlux*w / -patchFlux*(1 - w), diffusion
            # pGamma*(-delta) / -pGamma*delta; the glue's boundary list carries zeros for these faces
            cGamma = o.smul(self.nu, self.cMagSf)
            cwConv = capi.fv_limited_weights(self.ctx, self.cphi) if upwind else self.cw
            ci = o.sub(o.mul(self.cphi, cwConv), o.mul(cGamma, o.neg(self.cDelta)))
            cb = o.sub(o.mul(o.neg(self.cphi), o.rsub(1.0, cwConv)), o.mul(o.neg(cGamma), self.cDelta))
            ic, bc = cat([ic, self.zeroC3]), cat([bc, self.zeroC3])
        self.matU.set(diag, upper, lower, cb, ci)
        pnfU = (lambda f: capi.fv_patch_neighbour_field(a, 3, f)) if nC else (lambda f: None)
        perfs = {}
        if momentumPredictor:
            # solve(UEqn == -fvc::grad(p))
            src = o.add(source, o.mul(self.V, o.neg(self.grad(self.p)), 1, 3))
            UEqn = capi.FvMatrix(self.matU, 3, diag, src, self.U, self.V, ic, bc)
            perfs["U"] = UEqn.solve(USolver[0], USolver[1], pnf=pnfU(self.U), **(UControls or dict(tolerance=1e-5, relTol=0.0)))
        cont = []
        for corr in range(nCorr):
            UEqn = capi.FvMatrix(self.matU, 3, diag, source, self.U, self.V, ic, bc)
            rAU = o.rdiv(1.0, UEqn.A())
            HbyA = o.mul(rAU, UEqn.H(pnfU(self.U)), 1, 3)
            # phiHbyA = (interpolate(HbyA) & Sf) + interpolate(rAU)*ddtCorr(U, phi)
            phiCorr = o.sub(phi0, capi.fv_flux_linear(a, self.Sf, self.w, U0))
            coeff = o.rsub(1.0, o.smin(o.div(o.mag(phiCorr), o.sadd(o.mag(phi0), SMALL)), 1.0))
            ddtCorr = o.mul(o.smul(rDeltaT, coeff), phiCorr)
            rAUf = capi.fv_interpolate_linear(a, 1, self.w, rAU)
            phiHbyA = o.add(capi.fv_flux_linear(a, self.Sf, self.w, HbyA), o.mul(rAUf, ddtCorr))
            bphiHbyA = self.bUSf
            cphiHbyA, crAUf = self.cphi, None
            if nC:      # the same expression on the processor faces (their coupling coefficient is not zeroed)
                cphiCorr = o.sub(cphi0, o.dot3(self.cSf, self.interpolate_coupled(U0, 3)))
                ccoeff = o.rsub(1.0, o.smin(o.div(o.mag(cphiCorr), o.sadd(o.mag(cphi0), SMALL)), 1.0))
                crAUf = self.interpolate_coupled(rAU, 1)
                cphiHbyA = o.add(o.dot3(self.interpolate_coupled(HbyA, 3), self.cSf),
                                 o.mul(crAUf, o.mul(o.smul(rDeltaT, ccoeff), cphiCorr)))
            for nonOrth in range(nNonOrthCorr + 1):
                # pEqn: fvm::laplacian(rAU, p) == fvc::div(phiHbyA)
                pUpper, pDiag = capi.fv_laplacian_fill(a, self.delta, o.mul(rAUf, self.magSf))
                pSource = o.mul(self.V, self.div(phiHbyA, bphiHbyA, cphiHbyA))
                pCi = pCb = None
                if nC:
                    pGamma = o.mul(crAUf, self.cMagSf)
                    pCi, pCb = o.mul(pGamma, o.neg(self.cDelta)), o.mul(o.neg(pGamma), self.cDelta)
                pEqn = capi.FvMatrix(self.matP, 1, pDiag, pSource, self.p, self.V, self.zeroB1, self.zeroB1)
                pEqn.setReference(self.pRefCell, self.pRefValue)   # in place on pDiag / pSource, before the matrix copies them
                self.matP.set(pDiag, pUpper, None, pCb, pCi)
                perfs.setdefault("p", []).extend(pEqn.solve(pSolver[0], pSolver[1], gamg, **(pControls or dict(tolerance=1e-6, relTol=0.0))))
                if nonOrth == nNonOrthCorr:
                    pnfP = capi.fv_patch_neighbour_field(a, 1, self.p) if nC else None
                    internal, boundary, coupled = pEqn.flux(self.nB + nC, nC, pnfP)
                    self.phi = o.sub(phiHbyA, internal)
                    self.bphi = o.sub(bphiHbyA, boundary[: self.nB])
                    if nC:
                        self.cphi = o.sub(cphiHbyA, coupled)
            contErr = self.div(self.phi, self.bphi)
            tot = self._gsum3((contErr.abs() * self.V).sum(), (contErr * self.V).sum(), self.V.sum())
            cont.append((self.deltaT * tot[0] / tot[2], self.deltaT * tot[1] / tot[2]))
            # U = HbyA - rAU*fvc::grad(p)
            self.U = o.sub(HbyA, o.mul(rAU, self.grad(self.p), 1, 3), 3, 3)
        return perfs, cont

    def simple_step(self, alphaU=0.7, alphaP=0.3, divScheme="upwind", UControls=None, pControls=None,
                    USolver=("PBiCG", "DILU"), pSolver=("PCG", "DIC"), gamg=None, nNonOrthCorr=0):
        """One SIMPLE iteration (simpleFoam/UEqn.H:1-17, pEqn.H:1-40; laminar, single domain): UEqn = fvm::div(phi, U) -
        fvm::laplacian(nu, U); UEqn.relax(alphaU); solve(UEqn == -grad p); fvm::laplacian(rAU, p) == fvc::div(phiHbyA);
        phi = phiHbyA - pEqn.flux(); p.relax(alphaP); U = HbyA - rAU*grad p.  divScheme "upwind" (weights pos(phi),
        upwind.H:120-123) or "linear"."""
        assert not self.nC, "simple_step is single-domain"
        capi, o, a = self.capi, self.ops, self.addr
        wConv = capi.fv_limited_weights(self.ctx, self.phi) if divScheme == "upwind" else self.w
        cLower, cUpper, cDiag = capi.fv_convection_fill(a, wConv, self.phi)
        cBc = o.mul(o.neg(self.bphi), self.Ub, 1, 3)
        lUpper, lDiag = capi.fv_laplacian_fill(a, self.delta, o.smul(self.nu, self.magSf))
        diag = o.sub(cDiag, lDiag)
        upper = o.sub(cUpper, lUpper)
        lower = o.sub(cLower, lUpper)
        source = self._t(np.zeros((self.n, 3)))
        ic = o.sub(self.zeroB3, self.lIc)
        bc = o.sub(cBc, self.lBc)
        self.matU.set(diag, upper, lower)
        UEqn = capi.FvMatrix(self.matU, 3, diag, source, self.U, self.V, ic, bc)
        UEqn.relax(alphaU)                       # in place on diag / source (fvMatrix.C:1088-1345) ...
        self.matU.set(diag, upper, lower)        # ... so the matrix takes the relaxed diagonal
        perfs = {}
        src = o.add(source, o.mul(self.V, o.neg(self.grad(self.p)), 1, 3))
        UEqnP = capi.FvMatrix(self.matU, 3, diag, src, self.U, self.V, ic, bc)
        perfs["U"] = UEqnP.solve(USolver[0], USolver[1], **(UControls or dict(tolerance=1e-5, relTol=0.1)))
        rAU = o.rdiv(1.0, UEqn.A())
        HbyA = o.mul(rAU, UEqn.H(), 1, 3)
        phiHbyA = capi.fv_flux_linear(a, self.Sf, self.w, HbyA)
        bphiHbyA = self.bUSf
        rAUf = capi.fv_interpolate_linear(a, 1, self.w, rAU)
        pOld = self.p.clone()
        for nonOrth in range(nNonOrthCorr + 1):
            pUpper, pDiag = capi.fv_laplacian_fill(a, self.delta, o.mul(rAUf, self.magSf))
            pSource = o.mul(self.V, self.div(phiHbyA, bphiHbyA))
            pEqn = capi.FvMatrix(self.matP, 1, pDiag, pSource, self.p, self.V, self.zeroB1, self.zeroB1)
            pEqn.setReference(self.pRefCell, self.pRefValue)
            self.matP.set(pDiag, pUpper)
            perfs.setdefault("p", []).extend(pEqn.solve(pSolver[0], pSolver[1], gamg, **(pControls or dict(tolerance=1e-6, relTol=0.05))))
            if nonOrth == nNonOrthCorr:
                internal, boundary, _ = pEqn.flux(self.nB)
                self.phi = o.sub(phiHbyA, internal)
                self.bphi = o.sub(bphiHbyA, boundary[: self.nB])
        contErr = self.div(self.phi, self.bphi)
        tot = self._gsum3((contErr.abs() * self.V).sum(), (contErr * self.V).sum(), self.V.sum())
        cont = (tot[0] / tot[2], tot[1] / tot[2])
        self.p = o.add(pOld, o.smul(alphaP, o.sub(self.p, pOld)))     # p.relax()
        self.U = o.sub(HbyA, o.mul(rAU, self.grad(self.p), 1, 3), 3, 3)
        return perfs, cont

    def close(self):
        self.matU.close()
        self.matP.close()
        self.addr.close()


def cavity_rank(capi, ctx, torch, n, nRanks, rank, allsum, nu=0.01, deltaT=None, lid=(1.0, 0.0, 0.0)):
    """rank `rank` of the brick decomposition of the cavity (mesh.decompose): fixedValue walls + processor patches;
    `allsum(ndarray) -> ndarray` sums over the ranks (continuity report).  The pressure reference cell is global cell 0."""
    meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
    m = meshmod.decompose(n, nRanks, rank)
    walls, procs = m.wall_patches(), m.coupled_patches()
    bfc = np.concatenate([p.faceCells for p in walls]).astype(np.int32)
    bSf = np.concatenate([p.Sf for p in walls])
    Ub = np.concatenate([np.tile(lid if p.name == "movingWall" else (0.0, 0.0, 0.0), (len(p.faceCells), 1)) for p in walls])
    nB = len(bfc)
    ps, fc = m.patch_start_facecells()
    nC = len(fc)
    deltaT = deltaT if deltaT is not None else 0.5 * m.h / max(abs(v) for v in lid)
    ref = np.nonzero(m.cellGlobal == 0)[0]
    return m, IcoFoam(capi, ctx, torch, m.nCells, m.lower, m.upper, m.Sf(), m.magSf(), m.weights(), m.deltaCoeffs(),
                      m.volumes(), bfc, bSf, np.full(nB, m.h * m.h), np.full(nB, 2.0 / m.h), Ub, nu, deltaT,
                      pRefCell=int(ref[0]) if len(ref) else -1, cellCentres=m.cell_centres(), couPatchStart=ps,
                      couFaceCells=fc, neighbRank=[p.neighbRank for p in procs], couSf=np.concatenate([p.Sf for p in procs]),
                      couMagSf=np.full(nC, m.h * m.h), couWeights=np.full(nC, 0.5), couDeltaCoeffs=np.full(nC, 1.0 / m.h),
                      allsum=allsum)


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


# ---------------------------------------------------------------------------------------------------------
# icoFoam on a case directory (SURVEY.md section 8(f) ranks 2 + 3): constant/polyMesh, constant/transportProperties,
# system/controlDict, system/fvSolution and the start-time U, p as the reference's application reads them
# (createFields.H, readPISOControls.H, createTime.H).  Supported: 3-D single-domain cases whose boundary patches are all
# `fixedValue` for U and `zeroGradient` for p (the lid-driven cavity); anything else is refused by name.
# ---------------------------------------------------------------------------------------------------------
def _last_number(v):
    """`nu nu [0 2 -1 0 0 0 0] 0.01;` / `nu [..] 0.01;` / `0.01` -> 0.01"""
    if isinstance(v, (int, float)):
        return float(v)
    items = v if isinstance(v, (list, tuple)) else [v]
    nums = [x for x in items if isinstance(x, (int, float)) and not isinstance(x, bool)]
    if not nums:
        raise ValueError(f"no number in {v!r}")
    return float(nums[-1])


def perf_line(pf, fieldName):
    """solverPerformance::print (SolverPerformance.C:96-123)"""
    nm = pf.solverName.decode() if isinstance(pf.solverName, bytes) else str(pf.solverName)
    if pf.singular:
        return f"{nm}:  Solving for {fieldName}:  solution singularity"
    return (f"{nm}:  Solving for {fieldName}, Initial residual = {pf.initialResidual:g}, "
            f"Final residual = {pf.finalResidual:g}, No Iterations {pf.nIterations}")


def _time_name(t):
    s = f"{t:.6g}"      # Time::timeName with the default precision 6
    return s


def load_case(capi, ctx, torch, caseDir, rank=None, allsum=None):
    """Build the IcoFoam object and the run controls from a case directory.  rank = None: the undecomposed case;
    otherwise rank `rank` of a decomposed one: mesh and fields from <caseDir>/processor<rank>/ (decomposePar layout,
    `processor` patches last), dictionaries from the case itself; needs the communicator on `ctx` and
    `allsum(ndarray) -> ndarray` for the continuity report."""
    import os
    ff = importlib.import_module("rapidcfd-dev_b200.foamfile")
    j = lambda *a: os.path.join(caseDir, *a)
    base = caseDir if rank is None else j(f"processor{rank}")
    jb = lambda *a: os.path.join(base, *a)
    control = ff.read_dict(j("system", "controlDict"))
    fvSolution = ff.read_dict(j("system", "fvSolution"))
    transport = ff.read_dict(j("constant", "transportProperties"))
    pm = ff.read_poly_mesh(jb("constant", "polyMesh"))
    if pm.points is None:
        raise ValueError("constant/polyMesh needs points and faces (geometry)")
    coupled = pm.coupled_patches()
    if any(p.type != "processor" for p in coupled):
        raise NotImplementedError("icoFoam step: cyclic / processorCyclic patches are not supported yet")
    if coupled and rank is None:
        raise ValueError("processor patches in an undecomposed case")
    geo = pm.fv_geometry()
    nI = pm.nInternalFaces
    lower, upper = pm.ldu()
    startTime = _time_name(float(control.lookupOrDefault("startTime", 0)))
    n = pm.nCells
    Ufile = ff.read_field(jb(startTime, "U"), nInternal=n)
    pfile = ff.read_field(jb(startTime, "p"), nInternal=n)
    walls = [p for p in pm.patches if p.type != "processor"]
    if coupled and pm.patches[-len(coupled):] != coupled:
        raise ValueError("processor patches have to follow the physical patches")
    nW = sum(p.nFaces for p in walls)
    bfc = pm.owner[nI:nI + nW].astype(np.int32)
    Ub = np.zeros((nW, 3))
    for p in walls:
        s = slice(p.startFace - nI, p.startFace - nI + p.nFaces)
        ub, pb = Ufile["boundaryField"].get(p.name), pfile["boundaryField"].get(p.name)
        if ub is None or pb is None:
            raise KeyError(f"patch {p.name} has no entry in the boundaryField of U / p")
        if str(ub["type"]) != "fixedValue" or str(pb["type"]) != "zeroGradient":
            raise NotImplementedError(f"patch {p.name}: U {ub['type']} / p {pb['type']} -- this step supports "
                                      "fixedValue U with zeroGradient p only")
        if p.nFaces:
            Ub[s] = np.asarray(ub["value"], float).reshape(-1, 3) if np.ndim(ub["value"]) == 2 else np.asarray(ub["value"], float)
    # boundary deltaCoeffs: 1/|Cf - C_owner| (surfaceInterpolation.C:300-340, fvPatch::delta)
    wf = slice(nI, nI + nW)
    bDelta = 1.0 / np.linalg.norm(geo["Cf"][wf] - geo["C"][bfc], axis=1)
    piso = fvSolution.subDict("PISO")
    pRefCell = int(piso.lookupOrDefault("pRefCell", 0))
    kw = {}
    if coupled:
        ps, fc, nr = pm.coupled_interface_arrays()
        cf = slice(nI + nW, nI + nW + len(fc))
        # the reference cell is a GLOBAL label: the rank holding it uses its local one (setRefCell.C:100-140)
        cpa = jb("constant", "polyMesh", "cellProcAddressing")
        if os.path.exists(cpa) or os.path.exists(cpa + ".gz"):
            mine = np.nonzero(ff.read_list(cpa, "label") == pRefCell)[0]
            pRefCell = int(mine[0]) if len(mine) else -1
        elif rank != 0:
            pRefCell = -1
        # coupled weights and deltaCoeffs need the neighbour cell centres (coupledFvPatch::makeWeights,
        # processorFvPatch::delta): one patchNeighbourField exchange of C at start-up
        addr = capi.LduAddressing(ctx, n, lower, upper, ps, fc, nr, geo["C"])
        Cd = torch.from_numpy(np.ascontiguousarray(geo["C"], dtype=np.float64).ravel()).to(ctx.device)
        Cn = capi.fv_patch_neighbour_field(addr, 3, Cd).cpu().numpy().reshape(-1, 3)
        Sfc, Cfc, Co = geo["Sf"][cf], geo["Cf"][cf], geo["C"][fc]
        dOwn = np.abs(np.einsum("ij,ij->i", Sfc, Cfc - Co))
        dNei = np.abs(np.einsum("ij,ij->i", Sfc, Cn - Cfc))
        kw = dict(addr=addr, couPatchStart=ps, couFaceCells=fc, neighbRank=nr, couSf=Sfc, couMagSf=geo["magSf"][cf],
                  couWeights=dNei / (dOwn + dNei), couDeltaCoeffs=1.0 / np.linalg.norm(Cn - Co, axis=1), allsum=allsum)
    case = IcoFoam(capi, ctx, torch, n, lower, upper, geo["Sf"][:nI], geo["magSf"][:nI], geo["weights"], geo["deltaCoeffs"],
                   geo["V"], bfc, geo["Sf"][wf], geo["magSf"][wf], bDelta, Ub, _last_number(transport.lookup("nu")),
                   float(control.lookup("deltaT")), pRefCell, float(piso.lookupOrDefault("pRefValue", 0.0)),
                   cellCentres=geo["C"], **kw)
    case.U = case._t(Ufile["internalField"])
    case.p = case._t(pfile["internalField"])
    case.phi = capi.fv_flux_linear(case.addr, case.Sf, case.w, case.U)      # createPhi.H
    if case.nC:
        case.cphi = case.ops.dot3(case.interpolate_coupled(case.U, 3), case.cSf)
    us, up_, uc = ff.solver_controls(fvSolution, "U")
    ps_, pp, pc = ff.solver_controls(fvSolution, "p")
    run = dict(startTime=float(control.lookupOrDefault("startTime", 0)), endTime=float(control.lookup("endTime")),
               deltaT=float(control.lookup("deltaT")), nCorr=int(piso.lookupOrDefault("nCorrectors", 2)),
               nNonOrthCorr=int(piso.lookupOrDefault("nNonOrthogonalCorrectors", 0)),
               momentumPredictor=bool(ff._switch(piso.lookupOrDefault("momentumPredictor", "yes"))),
               USolver=(us, up_), UControls=uc, pSolver=(ps_, pp), pControls=pc, patches=pm.patches, nInternalFaces=nI,
               Ufile=Ufile, pfile=pfile, polyMesh=pm, geometry=geo, base=base, divScheme=_schemes_of_the_step(ff, j))
    return case, run


def _schemes_of_the_step(ff, j):
    """system/fvSchemes against what this step discretises with (icoFoam.C:55-103): every look-up the application makes is made
    here with the reference's rules (FvSchemes), and a scheme the step does not implement is refused by name instead of being
    silently replaced.  Returns the interpolation scheme of div(phi,U).  A case without system/fvSchemes (the reference would not
    start) runs with the cavity tutorial's schemes."""
    import os
    path = j("system", "fvSchemes")
    if not (os.path.exists(path) or os.path.exists(path + ".gz")):
        return "linear"
    fs = ff.FvSchemes(ff.read_dict(path))

    def need(what, got, allowed):
        if [str(t) for t in got] not in allowed:
            raise NotImplementedError(f"fvSchemes: {what} {' '.join(str(t) for t in got)} -- this step implements "
                                      + " | ".join(" ".join(a) for a in allowed))

    need("ddt(U)", fs.ddt("ddt(U)"), [["Euler"]])
    for g in ("grad(p)", "grad(U)"):
        need(g, fs.grad(g), [["Gauss", "linear"]])
    for lap in ("laplacian(nu,U)", "laplacian((1|A(U)),p)"):
        need(lap, fs.laplacian(lap), [["Gauss", "linear", c] for c in ("corrected", "uncorrected", "orthogonal")])
    need("interpolate(HbyA)", fs.interpolation("interpolate(HbyA)"), [["linear"]])
    if not fs.fluxRequired("p"):
        raise ValueError("fvSchemes: fluxRequired has no entry for p (fvMatrix::flux() of the pressure equation needs it, fvMatrix.C:857-866)")
    name, _, _ = ff.convection_scheme(fs.div("div(phi,U)"))
    if name not in ("linear", "upwind"):
        raise NotImplementedError(f"fvSchemes: div(phi,U) Gauss {name} -- limited schemes are built for scalar fields only "
                                  "(b200ldu_fv_limiter); this step implements Gauss linear | Gauss upwind")
    return name


def run_case(capi, ctx, torch, caseDir, log=print, write=True, maxSteps=None, rank=None, allsum=None):
    """The time loop of icoFoam.C:48-110 on `caseDir` (or on processor<rank>/ of it, see load_case); prints the
    reference's log lines; writes U and p of the last time step into <time>/ of the (processor) directory when `write`.
    Returns (case, list of per-step (perfs, continuity errors))."""
    import os
    ff = importlib.import_module("rapidcfd-dev_b200.foamfile")
    case, run = load_case(capi, ctx, torch, caseDir, rank, allsum)
    gamg = None
    if run["pSolver"][0] == "GAMG":
        geo = run["geometry"]
        Sf = geo["Sf"][:run["nInternalFaces"]]
        mag = np.linalg.norm(Sf, axis=1)
        wts = np.linalg.norm(Sf / np.sqrt(mag)[:, None] * np.array([1.0, 1.01, 1.02]), axis=1)   # faceAreaPair weights
        gamg = capi.GamgAgglomeration(case.addr, wts, int(run["pControls"].get("nCellsInCoarsestLevel", 10)),
                                      int(run["pControls"].get("mergeLevels", 1)))
    nSteps = int(round((run["endTime"] - run["startTime"]) / run["deltaT"]))
    if maxSteps is not None:
        nSteps = min(nSteps, maxSteps)
    history, cumulative, t = [], 0.0, run["startTime"]
    log("\nStarting time loop\n")
    for _ in range(nSteps):
        t += run["deltaT"]
        log(f"Time = {_time_name(t)}\n")
        perfs, cont = case.step(run["nCorr"], run["nNonOrthCorr"], run["UControls"] or None, run["pControls"] or None,
                                run["momentumPredictor"], run["USolver"], run["pSolver"], gamg, run["divScheme"])
        for k, pf in enumerate(perfs.get("U", [])):
            log(perf_line(pf, "U" + "xyz"[k]))
        per = len(perfs["p"]) // max(run["nCorr"], 1)
        for c in range(run["nCorr"]):
            for pf in perfs["p"][c * per:(c + 1) * per]:
                log(perf_line(pf, "p"))
            cumulative += cont[c][1]
            log(f"time step continuity errors : sum local = {cont[c][0]:g}, global = {cont[c][1]:g}, "
                f"cumulative = {cumulative:g}")
        history.append((perfs, cont))
    if write and nSteps:
        tdir = os.path.join(run["base"], _time_name(t))
        os.makedirs(tdir, exist_ok=True)
        U = case.U.cpu().numpy().reshape(-1, 3)
        p = case.p.cpu().numpy()
        ff.write_field(os.path.join(tdir, "U"), "volVectorField", [0, 1, -1, 0, 0, 0, 0], U,
                       run["Ufile"]["boundaryField"], location=_time_name(t))
        ff.write_field(os.path.join(tdir, "p"), "volScalarField", [0, 2, -2, 0, 0, 0, 0], p,
                       run["pfile"]["boundaryField"], location=_time_name(t))
    log("End\n")
    if gamg is not None:
        gamg.close()
    return case, history
