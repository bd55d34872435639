/*
 * ldu_oracle_fv.c -- CPU restatement of the finite-volume face-sum loops that
 * assemble the fvMatrix.  TEST INFRASTRUCTURE ONLY (see ldu_oracle.h).
 * Paths relative to /root/reference/src/finiteVolume/ (abbreviated FV/).
 * Floating-point contract: every face contribution is formed (product rounded) and
 * then added/subtracted in the order owner faces (ascending), neighbour faces
 * (losort order), boundary faces (ascending boundary-face index), then /V.
 */
#include "ldu_oracle_internal.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* fvc::surfaceIntegrate / surfaceSum: FV/finiteVolume/fvc/fvcSurfaceIntegrate.C:41-97
 * (functor), :138-203 (driver), :264-360 (surfaceSum: neiSign=+1, no division).
 * Boundary faces of all patches are passed concatenated in patch order. */
void orc_surface_integrate(const orc_addr *a, int nComp, const double *ssf, int nBFaces,
                           const int *bFaceCells, const double *bssf, const double *V,
                           double *out, int divideByV, int neiSign)
{
    for (int c = 0; c < a->nCells; c++) {
        for (int k = 0; k < nComp; k++) {
            double s = 0.0;
            for (int f = a->ownerStart[c]; f < a->ownerStart[c + 1]; f++)
                s = s + ssf[(size_t)f * nComp + k];
            for (int j = a->losortStart[c]; j < a->losortStart[c + 1]; j++) {
                int f = a->losort[j];
                if (neiSign < 0)
                    s = s - ssf[(size_t)f * nComp + k];
                else
                    s = s + ssf[(size_t)f * nComp + k];
            }
            out[(size_t)c * nComp + k] = s;
        }
    }
    for (int bf = 0; bf < nBFaces; bf++) { /* :174-200 */
        int c = bFaceCells[bf];
        for (int k = 0; k < nComp; k++)
            out[(size_t)c * nComp + k] = out[(size_t)c * nComp + k] + bssf[(size_t)bf * nComp + k];
    }
    if (divideByV) /* :202 */
        for (int c = 0; c < a->nCells; c++)
            for (int k = 0; k < nComp; k++) out[(size_t)c * nComp + k] /= V[c];
}

/* gaussGrad::gradf: FV/finiteVolume/gradSchemes/gaussGrad/gaussGrad.C:34-139 (functors),
 * :143-242 (driver).  nComp = 1: out is a vector per cell (3);  nComp = 3: out is a
 * tensor per cell, T[i][j] = Sf[i]*ssf[j] (outer product), row-major (9). */
void orc_gauss_grad(const orc_addr *a, int nComp, const double *Sf, const double *ssf,
                    int nBFaces, const int *bFaceCells, const double *bSf, const double *bssf,
                    const double *V, double *out)
{
    int nOut = 3 * nComp;
    for (int c = 0; c < a->nCells; c++) {
        double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int f = a->ownerStart[c]; f < a->ownerStart[c + 1]; f++)
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < nComp; j++) {
                    double p = Sf[(size_t)f * 3 + i] * ssf[(size_t)f * nComp + j];
                    acc[i * nComp + j] = acc[i * nComp + j] + p;
                }
        for (int k = a->losortStart[c]; k < a->losortStart[c + 1]; k++) {
            int f = a->losort[k];
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < nComp; j++) {
                    double p = Sf[(size_t)f * 3 + i] * ssf[(size_t)f * nComp + j];
                    acc[i * nComp + j] = acc[i * nComp + j] - p;
                }
        }
        for (int q = 0; q < nOut; q++) out[(size_t)c * nOut + q] = acc[q];
    }
    for (int bf = 0; bf < nBFaces; bf++) { /* :208-236 */
        int c = bFaceCells[bf];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < nComp; j++) {
                double p = bSf[(size_t)bf * 3 + i] * bssf[(size_t)bf * nComp + j];
                out[(size_t)c * nOut + i * nComp + j] = out[(size_t)c * nOut + i * nComp + j] + p;
            }
    }
    for (int c = 0; c < a->nCells; c++) /* :238 */
        for (int q = 0; q < nOut; q++) out[(size_t)c * nOut + q] /= V[c];
}

/* gaussLaplacianScheme::fvmLaplacianUncorrected:
 * FV/finiteVolume/laplacianSchemes/gaussLaplacianScheme/gaussLaplacianScheme.C:63-64
 * upper = deltaCoeffs*gammaMagSf ; negSumDiag (symmetric, diag starts at 0). */
void orc_laplacian_fill(const orc_addr *a, const double *deltaCoeffs, const double *gammaMagSf,
                        double *upper, double *diag)
{
    for (int f = 0; f < a->nFaces; f++) upper[f] = deltaCoeffs[f] * gammaMagSf[f];
    memset(diag, 0, sizeof(double) * (size_t)a->nCells);
    orc_negSumDiag(a, upper, NULL, diag);
}

/* gaussConvectionScheme::fvmDiv:
 * FV/finiteVolume/convectionSchemes/gaussConvectionScheme/gaussConvectionScheme.C:95-97 */
void orc_convection_fill(const orc_addr *a, const double *weights, const double *phi,
                         double *lower, double *upper, double *diag)
{
    for (int f = 0; f < a->nFaces; f++) {
        lower[f] = (-weights[f]) * phi[f];
        upper[f] = lower[f] + phi[f];
    }
    memset(diag, 0, sizeof(double) * (size_t)a->nCells);
    orc_negSumDiag(a, upper, lower, diag);
}

/* linear surface interpolation of a cell field to internal faces:
 * FV/interpolation/surfaceInterpolation/surfaceInterpolationScheme/
 * surfaceInterpolationScheme.C:272-351 (the one-weight form that interpolate(vf) reaches, :376-400;
 * not the two-weight form of :159-262) -- sf = w*(vf[own] - vf[nei]) + vf[nei]
 * (next-row component) */
void orc_interpolate_linear(const orc_addr *a, int nComp, const double *w, const double *vf,
                            double *sf)
{
    for (int f = 0; f < a->nFaces; f++)
        for (int k = 0; k < nComp; k++) {
            double own = vf[(size_t)a->l[f] * nComp + k], nei = vf[(size_t)a->u[f] * nComp + k];
            double d = own - nei;
            double p = w[f] * d;
            sf[(size_t)f * nComp + k] = p + nei;
        }
}

/* fvMatrix::addBoundaryDiag: FV/fvMatrices/fvMatrix/fvMatrix.C:209-226 (scalar cmpt) */
void orc_add_boundary_diag(int nBFaces, const int *bFaceCells, const double *internalCoeffs,
                           double *diag)
{
    for (int bf = 0; bf < nBFaces; bf++)
        diag[bFaceCells[bf]] = diag[bFaceCells[bf]] + internalCoeffs[bf];
}

/* fvMatrix::addBoundarySource (non-coupled part): FV/fvMatrices/fvMatrix/fvMatrix.C:290-312 */
void orc_add_boundary_source(int nBFaces, const int *bFaceCells, const double *boundaryCoeffs,
                             double *source)
{
    for (int bf = 0; bf < nBFaces; bf++)
        source[bFaceCells[bf]] = source[bFaceCells[bf]] + boundaryCoeffs[bf];
}

/* snGradScheme::snGrad on the internal faces: FV/finiteVolume/snGradSchemes/snGradScheme/snGradScheme.C:101-160
 * (snGradFunctor: d*(vf[neighbour] - vf[owner])) */
void orc_sngrad(const orc_addr *a, int nComp, const double *deltaCoeffs, const double *vf, double *out)
{
    for (int f = 0; f < a->nFaces; f++)
        for (int k = 0; k < nComp; k++) {
            double d = vf[(size_t)a->u[f] * nComp + k] - vf[(size_t)a->l[f] * nComp + k];
            out[(size_t)f * nComp + k] = deltaCoeffs[f] * d;
        }
}
