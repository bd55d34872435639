/*
 * ldu_oracle_omp.c -- all-core (OpenMP, rows in parallel) variant of the oracle's PCG
 * for the CPU baseline / bench.py --impl reference.  TEST INFRASTRUCTURE ONLY.
 * Same numerics as orc_pcg (PCG.C:69-208 with none/diagonal/AINV) except that global
 * sums are per-thread partial sums combined in thread order.
 */
#include "ldu_oracle_internal.h"
#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

int orc_max_threads(void) { return omp_get_max_threads(); }

static inline double row_sum(const orc_addr *a, int c, double init, const double *U,
                             const double *L, const double *x)
{
    double out = init;
    for (int f = a->ownerStart[c]; f < a->ownerStart[c + 1]; f++) out = out + U[f] * x[a->u[f]];
    for (int k = a->losortStart[c]; k < a->losortStart[c + 1]; k++) {
        int f = a->losort[k];
        out = out + L[f] * x[a->l[f]];
    }
    return out;
}

void orc_amul_omp(const orc_matrix *m, const double *psi, double *Apsi, int nThreads)
{
    const orc_addr *a = m->a;
#pragma omp parallel for num_threads(nThreads) schedule(static)
    for (int c = 0; c < a->nCells; c++)
        Apsi[c] = row_sum(a, c, m->diag[c] * psi[c], m->upper, m->lower, psi);
}

int orc_pcg_omp(const orc_matrix *m, int pk, const orc_controls *c, double *psi,
                const double *source, orc_perf *perf, int nT)
{
    const orc_addr *a = m->a;
    int n = a->nCells;
    memset(perf, 0, sizeof(*perf));
    double *pA = (double *)calloc((size_t)n, sizeof(double));
    double *wA = (double *)calloc((size_t)n, sizeof(double));
    double *rA = (double *)calloc((size_t)n, sizeof(double));
    double *rD = (double *)calloc((size_t)n, sizeof(double));
    double *tmp = (double *)calloc((size_t)n, sizeof(double));
    double wArA = 1e20, wArAold;

    orc_amul_omp(m, psi, wA, nT);
    double s0 = 0, sp = 0;
#pragma omp parallel for num_threads(nT) schedule(static) reduction(+ : sp)
    for (int i = 0; i < n; i++) {
        rA[i] = source[i] - wA[i];
        rD[i] = 1.0 / m->diag[i];
        sp += psi[i];
    }
    /* sumA (lduMatrixATmul.C:345-395), rows in parallel; single-domain matrices only (no interfaces) */
#pragma omp parallel for num_threads(nT) schedule(static)
    for (int i = 0; i < n; i++) {
        double out = m->diag[i];
        for (int f = a->ownerStart[i]; f < a->ownerStart[i + 1]; f++) out = out + m->upper[f];
        for (int k = a->losortStart[i]; k < a->losortStart[i + 1]; k++) out = out + m->lower[a->losort[k]];
        tmp[i] = out;
    }
    double avg = sp / (double)n, nf = 0;
#pragma omp parallel for num_threads(nT) schedule(static) reduction(+ : nf, s0)
    for (int i = 0; i < n; i++) {
        double t = avg * tmp[i];
        nf += fabs(wA[i] - t) + fabs(source[i] - t);
        s0 += fabs(rA[i]);
    }
    nf += 1e-20;
    perf->normFactor = nf;
    perf->initialResidual = perf->finalResidual = s0 / nf;
    const double *U = m->upper, *L = m->lower;

    int conv = perf->finalResidual < c->tolerance ||
               (c->relTol > 1e-20 && perf->finalResidual < c->relTol * perf->initialResidual);
    if (c->minIter > 0 || !conv) {
        do {
            wArAold = wArA;
            double dot = 0;
            if (pk == 2) {
#pragma omp parallel for num_threads(nT) schedule(static) reduction(+ : dot)
                for (int i = 0; i < n; i++) {
                    double out = 0.0;
                    for (int f = a->ownerStart[i]; f < a->ownerStart[i + 1]; f++) {
                        int nb = a->u[f];
                        out = out + (U[f] * rD[nb]) * rA[nb];
                    }
                    for (int k = a->losortStart[i]; k < a->losortStart[i + 1]; k++) {
                        int f = a->losort[k];
                        int nb = a->l[f];
                        out = out + (L[f] * rD[nb]) * rA[nb];
                    }
                    wA[i] = rD[i] * (rA[i] - out);
                    dot += wA[i] * rA[i];
                }
            } else {
#pragma omp parallel for num_threads(nT) schedule(static) reduction(+ : dot)
                for (int i = 0; i < n; i++) {
                    wA[i] = pk == 1 ? rD[i] * rA[i] : rA[i];
                    dot += wA[i] * rA[i];
                }
            }
            wArA = dot;
            if (perf->nIterations == 0) {
#pragma omp parallel for num_threads(nT) schedule(static)
                for (int i = 0; i < n; i++) pA[i] = wA[i];
            } else {
                double beta = wArA / wArAold;
#pragma omp parallel for num_threads(nT) schedule(static)
                for (int i = 0; i < n; i++) pA[i] = fma(beta, pA[i], wA[i]);
            }
            double wApA = 0;
#pragma omp parallel for num_threads(nT) schedule(static) reduction(+ : wApA)
            for (int i = 0; i < n; i++) {
                wA[i] = row_sum(a, i, m->diag[i] * pA[i], U, L, pA);
                wApA += wA[i] * pA[i];
            }
            if (fabs(wApA) / nf < 1e-300) {
                perf->singular = 1;
                break;
            }
            double alpha = wArA / wApA, sm = 0;
#pragma omp parallel for num_threads(nT) schedule(static) reduction(+ : sm)
            for (int i = 0; i < n; i++) {
                psi[i] = fma(alpha, pA[i], psi[i]);
                rA[i] = fma(-alpha, wA[i], rA[i]);
                sm += fabs(rA[i]);
            }
            perf->finalResidual = sm / nf;
            conv = perf->finalResidual < c->tolerance ||
                   (c->relTol > 1e-20 && perf->finalResidual < c->relTol * perf->initialResidual);
        } while ((perf->nIterations++ < c->maxIter && !conv) || perf->nIterations < c->minIter);
        perf->converged = conv;
    }
    free(pA);
    free(wA);
    free(rA);
    free(rD);
    free(tmp);
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * "Stock CPU OpenFOAM" baseline (CPU-stock-serial): what an unmodified
 * OpenFOAM-2.3.x rank does for `solver PCG; preconditioner DIC;` -- classic face-loop Amul
 * (lduMatrixATmul.C of stock OpenFOAM: Apsi[u[f]] += lower[f]*psi[l[f]]; Apsi[l[f]] += upper[f]*
 * psi[u[f]]) and the true diagonal incomplete-Cholesky preconditioner (calcReciprocalD +
 * forward/backward face sweeps).  RapidCFD does NOT run this: it silently replaces DIC by AINV
 * (LDU/preconditioners/DICPreconditioner/DICPreconditioner.C:43-59 of the reference); the stock
 * algorithm is restated here from OpenFOAM-2.3.x only to time the CPU baseline the north star
 * asks for.  Serial by construction (the sweeps carry a dependency along the face order).
 * --------------------------------------------------------------------------------------- */
int orc_pcg_stock_dic(const orc_matrix *m, const orc_controls *c, double *psi, const double *source,
                      orc_perf *perf)
{
    const orc_addr *a = m->a;
    const int n = a->nCells, nf = a->nFaces;
    const int *l = a->l, *u = a->u;
    const double *upper = m->upper, *diag = m->diag;
    memset(perf, 0, sizeof(*perf));
    double *pA = (double *)calloc((size_t)n, sizeof(double)), *wA = (double *)calloc((size_t)n, sizeof(double));
    double *rA = (double *)calloc((size_t)n, sizeof(double)), *rD = (double *)calloc((size_t)n, sizeof(double));
#define STOCK_AMUL(x, y)                                                        \
    do {                                                                        \
        for (int i = 0; i < n; i++) (y)[i] = diag[i] * (x)[i];                  \
        for (int f = 0; f < nf; f++) {                                          \
            (y)[u[f]] += upper[f] * (x)[l[f]];                                  \
            (y)[l[f]] += upper[f] * (x)[u[f]];                                  \
        }                                                                       \
    } while (0)
    STOCK_AMUL(psi, wA);
    double sp = 0, nfac = 0, s0 = 0;
    for (int i = 0; i < n; i++) {
        rA[i] = source[i] - wA[i];
        sp += psi[i];
    }
    orc_sumA(m, pA);
    double avg = sp / n;
    for (int i = 0; i < n; i++) {
        double t = avg * pA[i];
        nfac += fabs(wA[i] - t) + fabs(source[i] - t);
        s0 += fabs(rA[i]);
    }
    nfac += 1e-20;
    perf->normFactor = nfac;
    perf->initialResidual = perf->finalResidual = s0 / nfac;
    /* calcReciprocalD */
    for (int i = 0; i < n; i++) rD[i] = diag[i];
    for (int f = 0; f < nf; f++) rD[u[f]] -= upper[f] * upper[f] / rD[l[f]];
    for (int i = 0; i < n; i++) rD[i] = 1.0 / rD[i];
    double wArA = 1e20, wArAold;
    int conv = perf->finalResidual < c->tolerance ||
               (c->relTol > 1e-20 && perf->finalResidual < c->relTol * perf->initialResidual);
    if (c->minIter > 0 || !conv) {
        do {
            wArAold = wArA;
            for (int i = 0; i < n; i++) wA[i] = rD[i] * rA[i];
            for (int f = 0; f < nf; f++) wA[u[f]] -= rD[u[f]] * upper[f] * wA[l[f]];
            for (int f = nf - 1; f >= 0; f--) wA[l[f]] -= rD[l[f]] * upper[f] * wA[u[f]];
            wArA = 0;
            for (int i = 0; i < n; i++) wArA += wA[i] * rA[i];
            if (perf->nIterations == 0)
                memcpy(pA, wA, sizeof(double) * (size_t)n);
            else {
                double beta = wArA / wArAold;
                for (int i = 0; i < n; i++) pA[i] = wA[i] + beta * pA[i];
            }
            STOCK_AMUL(pA, wA);
            double wApA = 0;
            for (int i = 0; i < n; i++) wApA += wA[i] * pA[i];
            if (fabs(wApA) / nfac < 1e-300) {
                perf->singular = 1;
                break;
            }
            double alpha = wArA / wApA, sm = 0;
            for (int i = 0; i < n; i++) {
                psi[i] += alpha * pA[i];
                rA[i] -= alpha * wA[i];
                sm += fabs(rA[i]);
            }
            perf->finalResidual = sm / nfac;
            conv = perf->finalResidual < c->tolerance ||
                   (c->relTol > 1e-20 && perf->finalResidual < c->relTol * perf->initialResidual);
        } while ((perf->nIterations++ < c->maxIter && !conv) || perf->nIterations < c->minIter);
        perf->converged = conv;
    }
#undef STOCK_AMUL
    free(pA);
    free(wA);
    free(rA);
    free(rD);
    return 0;
}
