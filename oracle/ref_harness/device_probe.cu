/*
 * device_probe.cu -- NOT part of any library: compiles the reference's lduMatrixATmul.C and functor
 * headers for sm_100a (nvcc defaults, -fmad=true) so that `cuobjdump -sass` shows which a*b+c the
 * DEVICE build of the reference fuses into DFMA.  `make -C oracle probe` writes the DMUL/DADD/DFMA
 * listing of every functor kernel to oracle/_ref/device_probe_sass.txt; DESIGN.md section 2 quotes it.
 */
#include "lduMatrix.H" /* shim */

#include "lduMatrixATmul.C" /* reference */

#include "AINVPreconditionerF.H"
#include "JacobiSmootherF.H"

int Foam::lduMatrixSolutionCache::favourSpeed = 0;

using namespace Foam;
__global__ void probeAINV(AINVPreconditionerFunctor<false, 3> f, double *out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f(i);
}
__global__ void probeJacobi(JacobiSmootherFunctor<false, 3> f, double *out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f(i);
}
__global__ void probeAmul(matrixMultiplyFunctor<false, 3> f, double *out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f(i);
}
/* the solvers' vector updates (lduMatrixSolverFunctors.H) */
#include "lduMatrixSolverFunctors.H"
__global__ void probeAxpy(wAPlusBetaPAFunctor f, rAMinusAlphaWAFunctor g, const double *a, double *b, double *c, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        b[i] = f(a[i], b[i]);
        c[i] = g(c[i], a[i]);
    }
}
