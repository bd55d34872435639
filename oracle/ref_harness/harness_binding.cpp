/*
 * harness_binding.cpp -- TEST INFRASTRUCTURE: compiles the reference-side binding of the product,
 * rapidcfd-dev_b200/foam/b200Solver.H, together with the reference's own solver sources (everything
 * harness_solvers.cpp includes by path from /root/reference: lduMatrixSolver.C with solver::New and its
 * run-time selection tables, PCG.C, PBiCG.C ...).  In the resulting library the reference's
 * lduMatrix::solver::New (lduMatrixSolver.C:43-140) finds `b200PCG` / `b200PBiCG` / ... in the tables next to
 * its own PCG / PBiCG and constructs a b200Solver, whose solve() runs in libb200ldu.so on the GPU.
 * ref_solve(...) of harness_solvers.cpp is the entry point for both, so a test can run
 *     solver PCG;     preconditioner DIC;      -> the reference's PCG.C on the host
 *     solver b200PCG; preconditioner DIC;      -> the CUDA library through the binding
 * on the same matrix and compare the returned solverPerformance.
 * The stand-in gpuField of the harness lives in host memory: B200_FIELDS_ON_HOST makes the binding stage
 * fields through cudaMemcpy (in RapidCFD they are device pointers already).
 */
#define B200_FIELDS_ON_HOST
#include "harness_solvers.cpp"

#include "b200Solver.H"

B200_SOLVER_REGISTRATION
