/*
 * harness_mules.cpp -- runs the REFERENCE'S OWN explicit MULES on the CPU.  TEST INFRASTRUCTURE ONLY.
 * Included by path from /root/reference (symlinks in oracle/_ref/inc_mules/):
 *   FV/fvMatrices/solvers/MULES/MULES.H, MULESTemplates.C (explicitSolve :36-78, limiter :381-745 with its functors
 *   :143-377, limit :748-813), MULESFunctors.H, CMULES.H, CMULESTemplates.C (correct, limiterCorr, limitCorr)
 *   OpenFOAM/primitives/one/one.H, oneI.H, zero/zero.H, zeroI.H, fields/Fields/oneField/*, zeroField/*,
 *   fields/FieldFields/oneFieldField/*, fields/GeometricFields/geometricOneField/*
 * against oracle/ref_harness/shim_mules/ (+ shim/foam_shim.h).  What the shim restates instead of including is listed
 * at the top of shim_mules/mules_shim.h.
 */
#define NoRepository
#include "mules_shim.h"

#include "geometricOneField.H"
#include "MULES.H"
#include "CMULES.H" /* CMULESTemplates.C: correct :35-75, limiterCorr :375-704 with its functors :156-372, limitCorr :706-761 */

using namespace Foam;

namespace
{
struct Case {
    fvMesh mesh;
    volScalarField psi, psi0, rho, rho0;
    DimensionedInternalField Sp, Su;
    Case(int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart, const int *losort, int nP,
         const int *patchStart, const int *bFaceCells, const int *sortCellsStart, const int *sortCells, const int *sortAddr,
         const int *sortStart, const double *V, double rDeltaT, const double *psi_, const double *psi0_, const double *psiB,
         const double *rho_, const double *rho0_, const double *Sp_, const double *Su_, int nCoupledPatches)
    {
        lduAddressing &a = mesh.addr_;
        a.nCells_ = n;
        a.lower_.view(l, nF);
        a.upper_.view(u, nF);
        a.ownerStart_.view(ownerStart, n + 1);
        a.losortStart_.view(losortStart, n + 1);
        a.losort_.view(losort, nF);
        const int nB = patchStart[nP];
        mesh.nInternalFaces_ = nF;
        mesh.nFaces_ = nF + nB;
        mesh.time_.deltaT_ = 1.0 / rDeltaT;
        mesh.V_.f_.view(V, n);
        mesh.patches_.resize((size_t)nP);
        mesh.patchStart_.assign(patchStart, patchStart + nP + 1);
        a.patchCellsV_.resize((size_t)nP);
        a.patchSortV_.resize((size_t)nP);
        a.patchSortStartV_.resize((size_t)nP);
        psi.boundary_.resize((size_t)nP);
        for (int p = 0; p < nP; p++) {
            const int s = patchStart[p], e = patchStart[p + 1], cs = sortCellsStart[p], nu = sortCellsStart[p + 1] - cs;
            mesh.patches_[(size_t)p].faceCells_.view(bFaceCells + s, e - s);
            a.patchCellsV_[(size_t)p].view(sortCells + cs, nu);
            a.patchSortV_[(size_t)p].view(sortAddr + s, e - s);
            a.patchSortStartV_[(size_t)p].view(sortStart + cs + p, nu + 1); /* nu + 1 entries per patch */
            psi.boundary_[(size_t)p].view(psiB + s, e - s);
            psi.boundary_[(size_t)p].coupled_ = p >= nP - nCoupledPatches; /* the trailing patches (processor patches come last) */
        }
        psi.mesh_ = &mesh;
        psi.view(psi_, n);
        psi0.view(psi0_, n);
        psi.old_ = &psi0;
        if (rho_) {
            rho.view(rho_, n);
            rho0.view(rho0_ ? rho0_ : rho_, n);
            rho.old_ = &rho0;
        }
        if (Sp_) Sp.f_.view(Sp_, n);
        if (Su_) Su.f_.view(Su_, n);
    }
    void surface(surfaceScalarField &f, const double *flat) const /* own storage, internal then the patches */
    {
        f.mesh_ = &mesh;
        f.setSize(mesh.nInternalFaces_);
        std::copy(flat, flat + mesh.nInternalFaces_, f.data());
        f.boundary_.resize(mesh.patches_.size());
        for (size_t p = 0; p < f.boundary_.size(); p++) {
            const int s = mesh.patchStart_[p], e = mesh.patchStart_[p + 1];
            f.boundary_[p].setSize(e - s);
            std::copy(flat + mesh.nInternalFaces_ + s, flat + mesh.nInternalFaces_ + e, f.boundary_[p].data());
        }
    }
    void flat(const surfaceScalarField &f, double *out) const
    {
        std::copy(f.data(), f.data() + f.size(), out);
        for (size_t p = 0; p < f.boundary_.size(); p++)
            std::copy(f.boundary_[p].data(), f.boundary_[p].data() + f.boundary_[p].size(),
                      out + mesh.nInternalFaces_ + mesh.patchStart_[p]);
    }
};

template <class Rho, class SpT, class SuT>
void run(int mode, Case &c, double rDeltaT, const Rho &rho, const SpT &Sp, const SuT &Su, const double *a, const double *b,
         double psiMax, double psiMin, int nIter, double *out, const double *lambda0)
{
    if (mode == 0) { /* limiter: a = phiBD, b = phiCorr, lambda0 = the starting limiter (null: 1) -> out = allLambda */
        surfaceScalarField phiBD, phiCorr;
        c.surface(phiBD, a);
        c.surface(phiCorr, b);
        scalargpuField allLambda(c.mesh.nFaces(), 1.0);
        if (lambda0) std::copy(lambda0, lambda0 + c.mesh.nFaces(), allLambda.data());
        MULES::limiter(allLambda, rDeltaT, rho, c.psi, phiBD, phiCorr, Sp, Su, psiMax, psiMin, nIter);
        std::copy(allLambda.data(), allLambda.data() + allLambda.size(), out);
    } else if (mode == 1) { /* limit: a = phi, b = phiPsi -> out = the limited phiPsi */
        surfaceScalarField phi, phiPsi;
        c.surface(phi, a);
        c.surface(phiPsi, b);
        MULES::limit(rDeltaT, rho, c.psi, phi, phiPsi, Sp, Su, psiMax, psiMin, nIter, false);
        c.flat(phiPsi, out);
    } else if (mode == 3) { /* limiterCorr: a = phi, b = phiCorr, lambda0 -> out = allLambda */
        surfaceScalarField phi, phiCorr;
        c.surface(phi, a);
        c.surface(phiCorr, b);
        scalargpuField allLambda(c.mesh.nFaces(), 1.0);
        if (lambda0) std::copy(lambda0, lambda0 + c.mesh.nFaces(), allLambda.data());
        MULES::limiterCorr(allLambda, rDeltaT, rho, c.psi, phi, phiCorr, Sp, Su, psiMax, psiMin, nIter);
        std::copy(allLambda.data(), allLambda.data() + allLambda.size(), out);
    } else if (mode == 4) { /* limitCorr: a = phi, b = phiCorr -> out = lambda*phiCorr */
        surfaceScalarField phi, phiCorr;
        c.surface(phi, a);
        c.surface(phiCorr, b);
        MULES::limitCorr(rDeltaT, rho, c.psi, phi, phiCorr, Sp, Su, psiMax, psiMin, nIter);
        c.flat(phiCorr, out);
    } else if (mode == 5) { /* correct: a = phi (unused by the reference), b = phiCorr; out holds psi on entry -> the corrected psi */
        surfaceScalarField phi, phiCorr;
        c.surface(phi, a);
        c.surface(phiCorr, b);
        volScalarField psi;
        psi.mesh_ = &c.mesh;
        psi.view(out, c.psi.size());
        MULES::correct(rDeltaT, rho, psi, phi, phiCorr, Sp, Su);
    } else { /* explicitSolve: a = phiPsi -> out = psi */
        surfaceScalarField phiPsi;
        c.surface(phiPsi, a);
        volScalarField psi;
        psi.mesh_ = &c.mesh;
        psi.view(out, c.psi.size());
        psi.old_ = &c.psi0;
        MULES::explicitSolve(rDeltaT, rho, psi, phiPsi, Sp, Su);
    }
}
} // namespace

extern "C" int ref_mules(int mode, int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart,
                         const int *losort, int nP, const int *patchStart, const int *bFaceCells, const int *sortCellsStart,
                         const int *sortCells, const int *sortAddr, const int *sortStart, const double *V, double rDeltaT,
                         const double *psi, const double *psi0, const double *psiB, const double *rho, const double *rho0,
                         const double *Sp, const double *Su, const double *a, const double *b, double psiMax, double psiMin,
                         int nIter, double *out, int nCoupledPatches, const double *lambda0, double extremaCoeff)
{
    /* nCoupledPatches: the trailing patches answer coupled() = true and psiB holds their patchNeighbourField(); the shim's
     * syncFaceList is a no-op, so a multi-domain run calls this with nIter = 1 per sweep and takes the minimum with the other
     * side's values in between (oracle/mules_oracle.py reference_ranks) */
    Case c(n, nF, l, u, ownerStart, losortStart, losort, nP, patchStart, bFaceCells, sortCellsStart, sortCells, sortAddr, sortStart,
           V, rDeltaT, psi, psi0, psiB, rho, rho0, Sp, Su, nCoupledPatches);
    c.mesh.solverDict_.extremaCoeff_ = extremaCoeff; /* MULEScontrols.lookupOrDefault("extremaCoeff", 0.0), CMULESTemplates.C:398-401 */
    try {
        if (rho && Sp)
            run(mode, c, rDeltaT, c.rho, c.Sp, c.Su, a, b, psiMax, psiMin, nIter, out, lambda0);
        else if (rho)
            run(mode, c, rDeltaT, c.rho, zeroField(), zeroField(), a, b, psiMax, psiMin, nIter, out, lambda0);
        else if (Sp)
            run(mode, c, rDeltaT, geometricOneField(), c.Sp, c.Su, a, b, psiMax, psiMin, nIter, out, lambda0);
        else
            run(mode, c, rDeltaT, geometricOneField(), zeroField(), zeroField(), a, b, psiMax, psiMin, nIter, out, lambda0);
    } catch (const std::runtime_error &) {
        return -1;
    }
    return 0;
}
