/*
 * harness_solvers.cpp -- runs the REFERENCE'S OWN Krylov solver loops on the CPU.  TEST
 * INFRASTRUCTURE ONLY.  Included by path from /root/reference (symlinks in oracle/_ref/inc_solvers/):
 *   LDU/solvers/PCG/PCG.C:69-208                 PCG::solve
 *   LDU/solvers/PBiCG/PBiCG.C:68-246             PBiCG::solve
 *   LDU/solvers/PBiCGStab/PBiCGStab.C:66-300     PBiCGStab::solve (with its yA/zA slip, :263-270)
 *   LDU/solvers/smoothSolver/smoothSolver.C:77-193   smoothSolver::solve
 *   LDU/smoothers/Jacobi/JacobiSmoother.C:39-148, JacobiSmootherF.H
 *   LDU/preconditioners/AINVPreconditioner/AINVPreconditioner.C, AINVPreconditionerF.H
 *   LDU/preconditioners/diagonalPreconditioner/diagonalPreconditioner.C
 *   LDU/preconditioners/noPreconditioner/noPreconditioner.C
 *   LDU/lduMatrix/lduMatrixSolver.C:43-236       lduMatrix::solver::New, readControls, normFactor
 *   LDU/solvers/diagonalSolver/diagonalSolver.C
 *   LDU/lduMatrix/lduMatrixATmul.C, lduMatrixFunctors.H, lduMatrixSolverFunctors.H,
 *   LDU/lduAddressing/lduAddressingFunctors.H
 * against oracle/ref_harness/shim_solvers/ (+ shim/).  What the shims restate instead of including is
 * listed at the top of shim_solvers/solver_shim.h.
 */
#include "lduMatrix.H" /* shim_solvers */

#include "lduMatrixATmul.C"
#include "lduMatrixSolver.C" /* solver::New (run-time selection), readControls, normFactor */
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
#include "PBiCGStab.C"
#include "JacobiSmoother.C"
#include "GaussSeidelSmoother.C"
#include "smoothSolver.C"

#include <cstring>

namespace Foam
{
int lduMatrixSolutionCache::favourSpeed = 0;
int lduMatrix::debug = 0;
const gpuField<scalar> &lduMatrixSolutionCache::first(label size) { return ScratchPool::get("first", size); }
const gpuField<scalar> &lduMatrixSolutionCache::second(label size) { return ScratchPool::get("second", size); }

const dictionary dictionary::null;
} // namespace Foam

using namespace Foam;

extern "C" {
/* solver: "PCG" | "PBiCG" | "PBiCGStab"; returns 0, or -1 unknown solver, -2 unknown preconditioner.
 * perf[0..4] = initialResidual, finalResidual, nIterations, converged, singular; name receives the
 * printed solver name (preconditioner + typeName). */
int ref_solve(const char *solverName, const char *precond, int favourSpeed, int n, int nF, const int *l, const int *u,
              const int *ownerStart, const int *losortStart, const int *losort, const double *dg, const double *up,
              const double *lo, double tolerance, double relTol, int maxIter, int minIter, double *psi_io,
              const double *source, double *perf, char *name, int nameCap, int nSweeps, double omega)
{
    lduAddressing addr;
    std::vector<label> ownerSort(nF);
    std::vector<scalar> ls(nF), us(nF);
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
    scalargpuField lower(lo ? lo : up, nF), upper(up, nF), diag(dg, n), lowerSort(ls.data(), nF), upperSort(us.data(), nF);
    lduMatrix m;
    m.addr_ = &addr;
    m.lowerPtr_ = lo ? &lower : nullptr;
    m.upperPtr_ = &upper;
    m.diagPtr_ = &diag;
    m.lowerSortPtr_ = &lowerSort;
    m.upperSortPtr_ = &upperSort;
    m.level_ = 0;
    m.coarsest_ = false;
    lduMatrixSolutionCache::favourSpeed = favourSpeed;

    dictionary d;
    d.preconditioner = precond;
    d.smoother = precond; // smoothSolver reads `smoother` where the Krylov solvers read `preconditioner`
    d.nSweeps = nSweeps;
    d.omega = omega;
    d.tolerance = tolerance;
    d.relTol = relTol;
    d.maxIter = maxIter;
    d.minIter = minIter;
    FieldField<gpuField, scalar> noCoeffs(0);
    lduInterfaceFieldPtrsList noInterfaces;
    d.solver = solverName;
    autoPtr<lduMatrix::solver> s;
    try { // the reference's own run-time selection (lduMatrixSolver.C:43-140)
        s = lduMatrix::solver::New("psi", m, noCoeffs, noCoeffs, noInterfaces, d);
    } catch (const std::runtime_error &) {
        return -1; // "Unknown (a)symmetric matrix solver"
    }
    scalargpuField psi(psi_io, n), src(source, n);
    solverPerformance sp;
    try { // preconditioner::New / smoother::New are called inside solve(): unknown names are fatal there
        sp = s->solve(psi, src, 0);
    } catch (const std::runtime_error &e) {
        if (!strncmp(e.what(), "FatalError: ", 12)) { // abort(FatalError) with a message: a plug-in solver's failure
            if (name && nameCap > 0) {
                strncpy(name, e.what() + 12, (size_t)nameCap - 1);
                name[nameCap - 1] = 0;
            }
            return -3;
        }
        return -2;
    }
    perf[0] = sp.initialResidual();
    perf[1] = sp.finalResidual();
    perf[2] = sp.nIterations();
    perf[3] = sp.converged();
    perf[4] = sp.singular();
    if (name && nameCap > 0) {
        strncpy(name, sp.solverName().c_str(), (size_t)nameCap - 1);
        name[nameCap - 1] = 0;
    }
    return 0;
}
}
