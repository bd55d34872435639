/*
 * device_probe.cu -- NOT part of any library: compiles the reference's lduMatrixATmul.C and functor
 * headers for sm_100a (nvcc defaults, -fmad=true) so that `cuobjdump -sass` shows which a*b+c the
 * DEVICE build of the reference fuses into DFMA.  `make -C oracle probe` writes the DMUL/DADD/DFMA
 * listing of every functor kernel to oracle/_ref/device_probe_sass.txt; DESIGN.md section 2 quotes it.
 */
#include "lduMatrix.H" /* shim */

#include "lduMatrixATmul.C"     /* reference */
#include "lduMatrixTemplates.C" /* reference: lduMatrixfaceHFunctor */

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
/* the generic (non-"fast") row functor as lduMatrix::H composes it (lduMatrixTemplates.C:62-80) */
typedef matrixCoeffsMultiplyFunctor<scalar, scalar, negateUnaryOperatorFunctor<scalar, scalar>> HFun;
__global__ void probeH(lduAddressingFunctor<scalar, HFun, HFun, sumOp<scalar>, sumOp<scalar>> f, const double *in,
                       double *out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f(i, in[i]);
}
/* ... and the same functor's "fast" sibling as residual / H1 use it with favourSpeedOverMemory 2 */
__global__ void probeFast(lduAddressingFastFunctor<HFun, HFun> f, const double *in, double *out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f(i, in[i]);
}
/* coupled-interface update (coupledFvPatchField.C:246-256) */
__global__ void probeInterface(lduAddressingPatchFunctor<scalar, matrixInterfaceFunctor<scalar, false>, sumOp<scalar>> f,
                               double *out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f(i, out[i]);
}
/* faceH (lduMatrixTemplates.C:38-48) */
__global__ void probeFaceH(const double *u, const double *pu, const double *l, const double *pl, double *out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    lduMatrixfaceHFunctor<scalar> f;
    if (i < n) out[i] = f(thrust::make_tuple(u[i], pu[i], l[i], pl[i]));
}
