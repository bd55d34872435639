// solver_steps.cuh -- device-side scalar logic shared by solvers.cu and gamg.cu
#pragma once
#include "comm.h"
#include "ops.cuh"
#include "solvers.h"

static constexpr double GREAT_ = 1e20;   // SolverPerformance.H:269-275
static constexpr double SMALL_ = 1e-20;
static constexpr double VSMALL_ = 1e-300;

// ---- scalar step: sum partials (fixed order) [+ all-reduce] + scalar logic -------
P2PRed comm_p2p_red(b200ldu_ctx *ctx); // comm.cu

template <int NRED, class G>
int scalar_step_on(Solve &S, double *partials, int nPartials, G g)
{
    b200ldu_ctx *ctx = S.ctx;
    if (ctx->nRanks == 1 || NRED == 0 || ctx->p2p) {
        P2PRed p = (ctx->nRanks > 1 && NRED > 0) ? comm_p2p_red(ctx) : P2PRed();
        scalar_kernel<NRED, true, G><<<1, 256, 0, ctx->stream>>>(partials, nPartials, S.sc, g, p);
        ctx->launches++;
    } else {
        scalar_kernel<NRED, false, G><<<1, 256, 0, ctx->stream>>>(partials, nPartials, S.sc, g, P2PRed());
        ctx->launches++;
        TRY(comm_allreduce_sum(ctx, S.sc->sum, NRED)); // device pointer arithmetic only
        scalar_kernel<0, true, G><<<1, 256, 0, ctx->stream>>>(partials, 0, S.sc, g, P2PRed());
        ctx->launches++;
    }
    KERNEL_CHECK();
    return B200LDU_OK;
}

__device__ __forceinline__ bool check_convergence(SolverScalars *sc)
{
    // SolverPerformance.C:74-85
    bool c = sc->finalResidual < sc->tolerance ||
             (sc->relTol > SMALL_ && sc->finalResidual < sc->relTol * sc->initialResidual);
    sc->converged = c ? 1 : 0;
    return c;
}

__device__ __forceinline__ void hist_put(SolverScalars *sc, double *hist, int i, double v)
{
    if (hist && i < sc->histCap) hist[i] = v;
}

// end of a Krylov iteration body: (nIterations++ < maxIter && !converged) || nIterations < minIter
__device__ __forceinline__ void end_of_body(SolverScalars *sc, double *hist, double sumMag)
{
    sc->finalResidual = sumMag / sc->normFactor;
    hist_put(sc, hist, sc->nIterations + 1, sc->finalResidual);
    bool conv = check_convergence(sc);
    int n = sc->nIterations;
    sc->nIterations = n + 1;
    bool cont = (n < sc->maxIter && !conv) || (n + 1 < sc->minIter);
    if (!cont) sc->stop = 1;
}

#define V2(p) reinterpret_cast<double2 *>(p)
#define CV2(p) reinterpret_cast<const double2 *>(p)


template <int NRED, class G>
int scalar_step(Solve &S, int nPartials, G g)
{
    return scalar_step_on<NRED>(S, S.partials, nPartials, g);
}
