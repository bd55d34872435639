// mules_kernels.cuh -- device code of csrc/mules.cu (explicit MULES limiter), free of launch syntax so that
// tests/host_kernels/ can run it on the host against the oracle.
//
// Reference: FV/fvMatrices/solvers/MULES/CMULESTemplates.C:375-704 (MULES::limiterCorr, functors :156-372: `corr` below) and
// FV/fvMatrices/solvers/MULES/MULESTemplates.C:381-745 (MULES::limiter) with its functors
// (limiterMULESFunctor :143-258, patchMinMaxMULESFunctor :260-348, patchLambdaPfMULESFunctor :350-377,
// MULESFunctors.H: sumlPhiMULESFunctor, patchSumlPhiMULESFunctor, sumlPhipFinalMULESFunctor,
// lambdaIfMULESFunctor).  The reference runs one functor over the cells and then one per patch over the
// patch's cells; here one thread per cell walks its owner faces, its neighbour faces (losort) and its boundary
// faces (flat list in patch order) in that same order, so every sum has the reference's order.  Static mesh.  Coupled
// (processor / cyclic) patch faces are the trailing faces of the boundary list, with psiB = patchNeighbourField().
#ifndef B200LDU_MULES_KERNELS_CUH
#define B200LDU_MULES_KERNELS_CUH
#include <cstddef>

namespace mulesk
{
namespace
{
constexpr double MULES_VSMALL = 1e-300; // doubleScalar.H
constexpr double MULES_SMALL = 1e-15;

// Step 1: local extrema of psi over the face neighbours, sums of the bounded flux and of the positive /
// negative anti-diffusive fluxes, then the bounds turned into the flux budget of the cell
//   psiMaxn = V*((rho*rDeltaT - Sp)*min(psiMaxn, psiMax) - Su - (rho0*rDeltaT)*psi0) + sumPhiBD
//   psiMinn = V*(Su - (rho*rDeltaT - Sp)*max(psiMinn, psiMin) + (rho0*rDeltaT)*psi0) - sumPhiBD
// (MULESTemplates.C:445-560).  rho / rho0 / Sp / Su may be null: geometricOneField / zeroField (x*1 and x-0 are
// exact, so skipping the operation gives the bits of the reference's specialised operators).
__global__ void mules_bounds_kernel(int nCells, const int *__restrict__ ownerStart, const int *__restrict__ upper,
                                    const int *__restrict__ losortStart, const int *__restrict__ losort,
                                    const int *__restrict__ lower, const int *__restrict__ bStart,
                                    const int *__restrict__ bFaces, const double *__restrict__ psi,
                                    const double *__restrict__ psiB, const double *__restrict__ phiBD,
                                    const double *__restrict__ phiBDB, const double *__restrict__ phiCorr,
                                    const double *__restrict__ phiCorrB, const double *__restrict__ psi0,
                                    const double *__restrict__ rho, const double *__restrict__ rho0,
                                    const double *__restrict__ Sp, const double *__restrict__ Su,
                                    const double *__restrict__ V, double rDeltaT, double psiMaxG, double psiMinG,
                                    double *__restrict__ psiMaxn, double *__restrict__ psiMinn,
                                    double *__restrict__ sumPhip, double *__restrict__ mSumPhim, int corr, double extrema)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    double pMax = psiMinG, pMin = psiMaxG; // the search for the maximum starts from the global minimum (:445-446)
    double sumBD = 0, sp = MULES_VSMALL, sm = MULES_VSMALL;
    for (int f = ownerStart[c]; f < ownerStart[c + 1]; f++) {
        const double pn = psi[upper[f]];
        pMax = fmax(pMax, pn);
        pMin = fmin(pMin, pn);
        if (!corr) sumBD = __dadd_rn(sumBD, phiBD[f]);
        const double pc = phiCorr[f];
        if (pc > 0.0)
            sp = __dadd_rn(sp, pc);
        else
            sm = __dsub_rn(sm, pc);
    }
    for (int k = losortStart[c]; k < losortStart[c + 1]; k++) {
        const int f = losort[k];
        const double pn = psi[lower[f]];
        pMax = fmax(pMax, pn);
        pMin = fmin(pMin, pn);
        if (!corr) sumBD = __dsub_rn(sumBD, phiBD[f]);
        const double pc = phiCorr[f];
        if (pc > 0.0)
            sm = __dadd_rn(sm, pc);
        else
            sp = __dsub_rn(sp, pc);
    }
    if (bStart)
        for (int k = bStart[c]; k < bStart[c + 1]; k++) {
            const int bf = bFaces[k];
            pMax = fmax(pMax, psiB[bf]);
            pMin = fmin(pMin, psiB[bf]);
            if (!corr) sumBD = __dadd_rn(sumBD, phiBDB[bf]);
            const double pc = phiCorrB[bf];
            if (pc > 0.0)
                sp = __dadd_rn(sp, pc);
            else
                sm = __dsub_rn(sm, pc);
        }
    if (corr) { // limiterCorr: the extrema widened by extremaCoeff*(psiMax - psiMin) (CMULESTemplates.C:497-498)
        pMax = __dadd_rn(pMax, extrema);
        pMin = __dsub_rn(pMin, extrema);
    }
    pMax = fmin(pMax, psiMaxG);
    pMin = fmax(pMin, psiMinG);
    // (rho*rDeltaT - Sp); limiter: (rho0*rDeltaT)*psi0, limiterCorr: (rho*psi)*rDeltaT -- one rounding per written operator
    double a = rho ? __dmul_rn(rho[c], rDeltaT) : rDeltaT;
    if (Sp) a = __dsub_rn(a, Sp[c]);
    const double b = corr ? __dmul_rn(rho ? __dmul_rn(rho[c], psi[c]) : psi[c], rDeltaT)
                          : __dmul_rn(rho0 ? __dmul_rn(rho0[c], rDeltaT) : rDeltaT, psi0[c]);
    const double v = V[c];
    double up = __dmul_rn(a, pMax);
    if (Su) up = __dsub_rn(up, Su[c]);
    up = __dsub_rn(up, b);
    double lo = __dsub_rn(Su ? Su[c] : 0.0, __dmul_rn(a, pMin));
    lo = __dadd_rn(lo, b);
    if (corr) {
        psiMaxn[c] = __dmul_rn(v, up);
        psiMinn[c] = __dmul_rn(v, lo);
    } else {
        psiMaxn[c] = __dadd_rn(__dmul_rn(v, up), sumBD);
        psiMinn[c] = __dsub_rn(__dmul_rn(v, lo), sumBD);
    }
    sumPhip[c] = sp;
    mSumPhim[c] = sm;
}

// Step 2 (per limiter iteration): sums of the limited anti-diffusive fluxes per cell, then the cell limiters
//   lambdam = max(min((sumlPhip + psiMaxn)/(mSumPhim - SMALL), 1), 0)
//   lambdap = max(min((mSumlPhim + psiMinn)/(sumPhip + SMALL), 1), 0)     (MULESTemplates.C:563-640)
__global__ void mules_cell_lambda_kernel(int nCells, const int *__restrict__ ownerStart, const int *__restrict__ losortStart,
                                         const int *__restrict__ losort, const int *__restrict__ bStart,
                                         const int *__restrict__ bFaces, const double *__restrict__ lambda,
                                         const double *__restrict__ lambdaB, const double *__restrict__ phiCorr,
                                         const double *__restrict__ phiCorrB, const double *__restrict__ psiMaxn,
                                         const double *__restrict__ psiMinn, const double *__restrict__ sumPhip,
                                         const double *__restrict__ mSumPhim, double *__restrict__ lambdam,
                                         double *__restrict__ lambdap)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    double slp = 0.0, mslm = 0.0;
    for (int f = ownerStart[c]; f < ownerStart[c + 1]; f++) {
        const double lp = __dmul_rn(lambda[f], phiCorr[f]);
        if (lp > 0.0)
            slp = __dadd_rn(slp, lp);
        else
            mslm = __dsub_rn(mslm, lp);
    }
    for (int k = losortStart[c]; k < losortStart[c + 1]; k++) {
        const int f = losort[k];
        const double lp = __dmul_rn(lambda[f], phiCorr[f]);
        if (lp > 0.0)
            mslm = __dadd_rn(mslm, lp);
        else
            slp = __dsub_rn(slp, lp);
    }
    if (bStart)
        for (int k = bStart[c]; k < bStart[c + 1]; k++) {
            const int bf = bFaces[k];
            const double lp = __dmul_rn(lambdaB[bf], phiCorrB[bf]);
            if (lp > 0.0)
                slp = __dadd_rn(slp, lp);
            else
                mslm = __dsub_rn(mslm, lp);
        }
    lambdam[c] = fmax(fmin(__ddiv_rn(__dadd_rn(slp, psiMaxn[c]), __dsub_rn(mSumPhim[c], MULES_SMALL)), 1.0), 0.0);
    lambdap[c] = fmax(fmin(__ddiv_rn(__dadd_rn(mslm, psiMinn[c]), __dadd_rn(sumPhip[c], MULES_SMALL)), 1.0), 0.0);
}

// Step 3: face limiters from the cell limiters (lambdaIfMULESFunctor; boundary: patchLambdaPfMULESFunctor, outflow
// faces only; the trailing nCoupled boundary faces are coupled patch faces: coupledPatchLambdaPfMULESFunctor, every face).
// i < nFaces: internal face i; else boundary face i - nFaces.
__global__ void mules_face_lambda_kernel(int nFaces, int nBFaces, int nCoupled, int corr, const int *__restrict__ lower, const int *__restrict__ upper,
                                         const int *__restrict__ bFaceCells, const double *__restrict__ phiCorr,
                                         const double *__restrict__ phiCorrB, const double *__restrict__ phiBDB,
                                         const double *__restrict__ lambdam, const double *__restrict__ lambdap,
                                         double *__restrict__ lambda, double *__restrict__ lambdaB)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nFaces) {
        const int o = lower[i], n = upper[i];
        const double l = lambda[i];
        lambda[i] = phiCorr[i] > 0.0 ? fmin(l, fmin(lambdap[o], lambdam[n])) : fmin(l, fmin(lambdam[o], lambdap[n]));
    } else if (i < nFaces + nBFaces) {
        const int bf = i - nFaces;
        const double l = lambdaB[bf], pc = phiCorrB[bf];
        // outflow faces only: limiter tests phiBD + phiCorr, limiterCorr the total flux (phiBDB then holds phi's boundary values)
        if (bf >= nBFaces - nCoupled || (corr ? phiBDB[bf] : __dadd_rn(phiBDB[bf], pc)) > MULES_SMALL * MULES_SMALL) {
            const int c = bFaceCells[bf];
            lambdaB[bf] = pc > 0.0 ? fmin(l, lambdap[c]) : fmin(l, lambdam[c]);
        }
    }
}

// syncTools::syncFaceList(mesh, allLambda, minOp<scalar>()) on the coupled faces (MULESTemplates.C:743): mine = min(mine, theirs)
__global__ void mules_sync_min_kernel(int n, double *__restrict__ mine, const double *__restrict__ theirs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mine[i] = fmin(mine[i], theirs[i]);
}
} // namespace
} // namespace mulesk
#endif
