/*
 * harness_gamgsolve.cpp -- runs the REFERENCE'S OWN GAMG cycle on the CPU.  TEST INFRASTRUCTURE ONLY.
 * Included by path from /root/reference (symlinks in oracle/_ref/inc_gamgsolve/):
 *   GAMG/GAMGSolverSolve.C:59-619        GAMGSolver::solve, Vcycle, initVcycle, solveCoarsestLevel
 *   GAMG/GAMGSolverScale.C:59-171        GAMGSolver::scale
 *   GAMG/GAMGSolverInterpolate.C:45-110  GAMGSolver::interpolate
 *   matrices/scalarMatrices/scalarMatrices.C:42-146, scalarMatricesTemplates.C:119-164   the coarsest-level LU
 *   + the smoother, Krylov solvers (coarsest level when directSolveCoarsest is off), preconditioners and
 *     lduMatrixATmul.C as in harness_solvers.cpp
 * against oracle/ref_harness/shim_gamgsolve/ (+ shim_solvers/, shim/).  The level hierarchy (restrict maps,
 * coarse addressing, coarse coefficients) is passed in by the test.
 */
#include "GAMGSolver.H"     /* shim */
#include "scalarMatrices.H" /* shim */

#include "scalarMatrices.C"          /* reference: LUDecompose (Crout, implicit pivoting) */
#include "scalarMatricesTemplates.C" /* reference: LUBacksubstitute */

namespace Foam
{
// LUscalarMatrix.C:47-56 / LUscalarMatrixTemplates.C:110-126 on one rank: decompose once, back-substitute per solve
class LUscalarMatrix
{
    scalarSquareMatrix lu_;
    labelList piv_;

public:
    LUscalarMatrix(label n, const std::vector<scalar> &dense) : lu_(n), piv_(n)
    {
        for (label i = 0; i < n; i++)
            for (label j = 0; j < n; j++) lu_[i][j] = dense[(size_t)i * n + j];
        LUDecompose(lu_, piv_);
    }
    void solve(scalarField &x) const
    {
        List<scalar> b(x.size());
        for (label i = 0; i < x.size(); i++) b[i] = x.data()[i];
        LUBacksubstitute(lu_, piv_, b);
        for (label i = 0; i < x.size(); i++) x.data()[i] = b[i];
    }
};
} // namespace Foam

#include "lduMatrixATmul.C"
#include "lduMatrixSolver.C" /* solver base: constructor, readControls, normFactor, New */
#include "diagonalSolver.C"
#include "lduMatrixPreconditioner.C" /* preconditioner::New / getName with the reference's tables */
#include "lduMatrixSmoother.C"       /* smoother::New / getName */
#include "AINVPreconditioner.C"
#include "DICPreconditioner.C"
#include "DILUPreconditioner.C"
#include "diagonalPreconditioner.C"
#include "noPreconditioner.C"
#include "PCG.C"
#include "PBiCG.C"
#include "JacobiSmoother.C"
#include "GaussSeidelSmoother.C"
#include "GAMGSolverSolve.C"
#include "GAMGSolverScale.C"
#include "GAMGSolverInterpolate.C"

#include <cmath>

namespace Foam
{
int lduMatrixSolutionCache::favourSpeed = 0;
int lduMatrix::debug = 0;
label UPstream::warnComm = -1;
defineTypeNameAndDebug(GAMGSolver, 0); // GAMGSolver.C:35
GAMGSolver::~GAMGSolver() { delete coarsestBufferPtr_; }
const gpuField<scalar> &lduMatrixSolutionCache::first(label size) { return ScratchPool::get("first", size); }
const gpuField<scalar> &lduMatrixSolutionCache::second(label size) { return ScratchPool::get("second", size); }

const dictionary dictionary::null;
} // namespace Foam

using namespace Foam;

namespace
{
struct Level { // storage behind one lduMatrix
    lduAddressing addr;
    scalargpuField lower, upper, diag, lowerSort, upperSort;
    std::vector<label> ownerSort;
    std::vector<scalar> ls, us;
    void fill(lduMatrix &m, int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart,
              const int *losort, const double *dg, const double *up, const double *lo, int level, bool coarsest)
    {
        ownerSort.resize(nF);
        ls.resize(nF);
        us.resize(nF);
        for (int k = 0; k < nF; k++) {
            ownerSort[k] = l[losort[k]];
            ls[k] = (lo ? lo : up)[losort[k]];
            us[k] = up[losort[k]];
        }
        addr.nCells_ = n;
        addr.lower_.view(l, nF);
        addr.upper_.view(u, nF);
        addr.ownerStart_.view(ownerStart, n + 1);
        addr.losortStart_.view(losortStart, n + 1);
        addr.losort_.view(losort, nF);
        addr.ownerSort_.view(ownerSort.data(), nF);
        lower.view(lo ? lo : up, nF);
        upper.view(up, nF);
        diag.view(dg, n);
        lowerSort.view(ls.data(), nF);
        upperSort.view(us.data(), nF);
        m.addr_ = &addr;
        m.lowerPtr_ = lo ? &lower : nullptr;
        m.upperPtr_ = &upper;
        m.diagPtr_ = &diag;
        m.lowerSortPtr_ = &lowerSort;
        m.upperSortPtr_ = &upperSort;
        m.level_ = level;
        m.coarsest_ = coarsest;
    }
};
} // namespace

extern "C" {
/* levels: arrays indexed 0 = finest matrix, 1..nLevels = coarse levels.  restrictMaps[k] (k = 0..nLevels-1) maps
 * the cells of level k to level k+1.  ctl: nPreSweeps, preSweepsLevelMultiplier, maxPreSweeps, nPostSweeps,
 * postSweepsLevelMultiplier, maxPostSweeps, nFinestSweeps, interpolateCorrection, scaleCorrection,
 * directSolveCoarsest, maxIter, minIter, favourSpeed.  perf: as ref_solve.  Returns 0, or -3 if the reference
 * code path ends in notImplemented(). */
int ref_gamg_solve(int nLevels, const int *nCells, const int *nFaces, const int *const *l, const int *const *u,
                   const int *const *ownerStart, const int *const *losortStart, const int *const *losort,
                   const double *const *diag, const double *const *upper, const double *const *lower,
                   const int *const *restrictMaps, const int *ctl, double tolerance, double relTol, double omega,
                   double *psi_io, const double *source, double *perf)
{
    std::vector<Level> store((size_t)nLevels + 1);
    lduMatrix fine;
    store[0].fill(fine, nCells[0], nFaces[0], l[0], u[0], ownerStart[0], losortStart[0], losort[0], diag[0], upper[0],
                  lower[0], 0, false);
    GAMGAgglomeration agg;
    for (int k = 0; k < nLevels; k++) {
        agg.restrictAddr.emplace_back(restrictMaps[k], restrictMaps[k] + nCells[k]);
        agg.nCoarse.push_back(nCells[k + 1]);
    }
    dictionary d;
    d.smoother = "GaussSeidel";
    d.omega = omega;
    d.tolerance = tolerance;
    d.relTol = relTol;
    d.maxIter = ctl[10];
    d.minIter = ctl[11];
    lduMatrixSolutionCache::favourSpeed = ctl[12];
    FieldField<gpuField, scalar> noCoeffs(0);
    lduInterfaceFieldPtrsList noInterfaces;
    GAMGSolver g("psi", fine, noCoeffs, noCoeffs, noInterfaces, d, agg);
    g.nPreSweeps_ = ctl[0];
    g.preSweepsLevelMultiplier_ = ctl[1];
    g.maxPreSweeps_ = ctl[2];
    g.nPostSweeps_ = ctl[3];
    g.postSweepsLevelMultiplier_ = ctl[4];
    g.maxPostSweeps_ = ctl[5];
    g.nFinestSweeps_ = ctl[6];
    g.interpolateCorrection_ = ctl[7] != 0;
    g.scaleCorrection_ = ctl[8] != 0;
    g.directSolveCoarsest_ = ctl[9] != 0;
    g.matrixLevels_.setSize(nLevels);
    g.interfaceLevels_.setSize(nLevels);
    g.interfaceLevelsBouCoeffs_.setSize(nLevels);
    g.interfaceLevelsIntCoeffs_.setSize(nLevels);
    for (int k = 0; k < nLevels; k++) {
        lduMatrix *m = new lduMatrix();
        store[(size_t)k + 1].fill(*m, nCells[k + 1], nFaces[k + 1], l[k + 1], u[k + 1], ownerStart[k + 1],
                                  losortStart[k + 1], losort[k + 1], diag[k + 1], upper[k + 1], lower[k + 1], k + 1,
                                  k == nLevels - 1);
        g.matrixLevels_.set(k, m);
        g.interfaceLevels_.set(k, new lduInterfaceFieldPtrsList());
        g.interfaceLevelsBouCoeffs_.set(k, new FieldField<gpuField, scalar>(0));
        g.interfaceLevelsIntCoeffs_.set(k, new FieldField<gpuField, scalar>(0));
    }
    if (g.directSolveCoarsest_) { // GAMGSolver.C:146-172: LU of the coarsest matrix
        const int c = nLevels, n = nCells[c];
        std::vector<scalar> dense((size_t)n * n, 0.0);
        for (int i = 0; i < n; i++) dense[(size_t)i * n + i] = diag[c][i];
        for (int f = 0; f < nFaces[c]; f++) {
            dense[(size_t)l[c][f] * n + u[c][f]] = upper[c][f];
            dense[(size_t)u[c][f] * n + l[c][f]] = (lower[c] ? lower[c] : upper[c])[f];
        }
        g.coarsestLUMatrixPtr_ = autoPtr<LUscalarMatrix>(new LUscalarMatrix(n, dense));
        g.coarsestBufferPtr_ = new scalarField(n);
    }
    scalargpuField psi(psi_io, nCells[0]), src(source, nCells[0]);
    try {
        solverPerformance sp = g.solve(psi, src, 0);
        perf[0] = sp.initialResidual();
        perf[1] = sp.finalResidual();
        perf[2] = sp.nIterations();
        perf[3] = sp.converged();
        perf[4] = sp.singular();
    } catch (const std::runtime_error &) {
        return -3;
    }
    return 0;
}
}
