/*
 * Host execution of the device code of csrc/fvmatrix.cu and csrc/fieldops.cu (their *_kernels.cuh headers): the CUDA
 * qualifiers are defined away, the rounding intrinsics become plain fp64 operations (compiled with
 * -ffp-contract=off, so each is one rounding as on the device) and a launcher loop sets blockIdx / threadIdx for
 * every thread of the grid in turn.  TEST INFRASTRUCTURE ONLY: lets the CPU suite run the same kernel source
 * against the oracle where no GPU is available.  It executes the kernels' index arithmetic, ordering and rounding;
 * it does not exercise the launch code of the .cu files nor anything about the real device.
 */
#include <cmath>
#include <cstddef>

#define __global__
#define __device__
#define __forceinline__ inline
struct Dim3 {
    unsigned x, y, z;
};
static thread_local Dim3 blockIdx, blockDim, threadIdx;
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
using std::fabs;
using std::fmax;
using std::fmin;

#include "fieldops_kernels.cuh"
#include "fvmatrix_kernels.cuh"
#include "mules_kernels.cuh"
#include "fv_kernels.cuh"

template <class K, class... A> static void launch(long long nThreads, int block, K k, A... args)
{
    const long long nBlocks = (nThreads + block - 1) / block;
    blockDim = {(unsigned)block, 1, 1};
    for (long long b = 0; b < nBlocks; b++)
        for (int t = 0; t < block; t++) {
            blockIdx = {(unsigned)b, 0, 0};
            threadIdx = {(unsigned)t, 0, 0};
            k(args...);
        }
}

using namespace fvmk;

extern "C" {
struct HostCase { /* what csrc/fvmatrix.cu reads from the addressing and the matrix handle */
    int nCells, nFaces, nB, nC;
    const int *l, *u, *ownerStart, *losortStart, *losort;
    const int *bStart, *bFaces, *bFaceCells, *cStart, *cFaces, *cFaceCells;
    const double *diag, *upper, *lower, *couInt, *couBou;
};

static BoundaryLists lists(const HostCase *h)
{
    return BoundaryLists{h->nB ? h->bStart : nullptr, h->bFaces, h->nC ? h->cStart : nullptr, h->cFaces};
}

void hk_boundary_diag(const HostCase *h, int nc, int cmpt, const double *ic, const double *in, double *out)
{
    launch(h->nCells, 256, boundary_diag_kernel, h->nCells, lists(h), ic, nc, cmpt, h->couInt, in, out);
}

void hk_boundary_source(const HostCase *h, int nc, const double *bc, const double *pnf, const double *in, double *out)
{
    if (nc == 1)
        launch(h->nCells, 256, boundary_source_kernel<1>, h->nCells, lists(h), bc, h->couBou, pnf, in, out);
    else
        launch(h->nCells, 256, boundary_source_kernel<3>, h->nCells, lists(h), bc, h->couBou, pnf, in, out);
}

void hk_component(const HostCase *h, int nc, int k, int withCoupled, const double *pnf, const double *in, double *out)
{
    BoundaryLists L = lists(h);
    if (!withCoupled) L.cStart = nullptr;
    launch(h->nCells, 256, component_kernel, h->nCells, nc, k, L, h->couBou, pnf, in, out);
}

void hk_set_component(const HostCase *h, int nc, int k, const double *in, double *out)
{
    launch(h->nCells, 256, set_component_kernel, h->nCells, nc, k, in, out);
}

void hk_A(const HostCase *h, int nc, const double *ic, const double *V, double *out)
{
    launch(h->nCells, 256, A_kernel, h->nCells, lists(h), ic, nc, h->couInt, h->diag, V, out);
}

void hk_H(const HostCase *h, int nc, const double *psi, const double *source, const double *bc, const double *pnf,
          const double *V, double *out)
{
    if (nc == 1)
        launch(h->nCells, 128, H_kernel<1>, h->nCells, h->ownerStart, h->u, h->losortStart, h->losort, h->l, h->upper,
               h->lower, lists(h), bc, h->couBou, pnf, psi, source, V, out);
    else
        launch(h->nCells, 128, H_kernel<3>, h->nCells, h->ownerStart, h->u, h->losortStart, h->losort, h->l, h->upper,
               h->lower, lists(h), bc, h->couBou, pnf, psi, source, V, out);
}

void hk_flux(const HostCase *h, int nc, const double *psi, const double *ic, const double *bc, const double *pnf,
             double *flux, double *bflux, double *cflux)
{
    if (nc == 1) {
        launch(h->nFaces, 256, flux_internal_kernel<1>, h->nFaces, h->l, h->u, h->upper, h->lower, psi, flux);
        if (h->nB) launch(h->nB, 256, flux_boundary_kernel<1>, h->nB, h->bFaceCells, ic, 1, bc, 1, (const double *)nullptr, psi, bflux);
        if (h->nC) launch(h->nC, 256, flux_boundary_kernel<1>, h->nC, h->cFaceCells, h->couInt, 1, h->couBou, 1, pnf, psi, cflux);
    } else {
        launch(h->nFaces, 256, flux_internal_kernel<3>, h->nFaces, h->l, h->u, h->upper, h->lower, psi, flux);
        if (h->nB) launch(h->nB, 256, flux_boundary_kernel<3>, h->nB, h->bFaceCells, ic, 3, bc, 3, (const double *)nullptr, psi, bflux);
        if (h->nC) launch(h->nC, 256, flux_boundary_kernel<3>, h->nC, h->cFaceCells, h->couInt, 1, h->couBou, 1, pnf, psi, cflux);
    }
}

void hk_residual_source(const HostCase *h, const double *ic, const double *psi, const double *source, double *out)
{
    launch(h->nCells, 256, residual_source_kernel, h->nCells, lists(h), ic, h->couInt, psi, source, out);
}

void hk_relax(const HostCase *h, int nc, double alpha, const double *psi, const double *ic, double *diag, double *source)
{
    if (nc == 1)
        launch(h->nCells, 128, relax_kernel<1>, h->nCells, h->ownerStart, h->losortStart, h->losort, h->upper, h->lower,
               lists(h), ic, h->couInt, h->couBou, alpha, psi, diag, source);
    else
        launch(h->nCells, 128, relax_kernel<3>, h->nCells, h->ownerStart, h->losortStart, h->losort, h->upper, h->lower,
               lists(h), ic, h->couInt, h->couBou, alpha, psi, diag, source);
}

void hk_set_reference(int cell, int nc, const double *v, double *diag, double *source)
{
    launch(1, 1, set_reference_kernel, cell, nc, v[0], nc > 1 ? v[1] : 0.0, nc > 2 ? v[2] : 0.0, diag, source);
}

void hk_field_binary(int op, long long n, int ncA, const double *a, int ncB, const double *b, double *out)
{
    const int nc = ncA > ncB ? ncA : ncB;
    launch(n * nc, 256, fieldk::binary_kernel, n, nc, ncA, ncB, op, a, b, out);
}

void hk_field_unary(int op, long long n, double s, const double *a, double *out)
{
    launch(n, 256, fieldk::unary_kernel, n, op, s, a, out);
}

void hk_field_dot3(long long n, const double *a, const double *b, double *out) { launch(n, 256, fieldk::dot3_kernel, n, a, b, out); }
void hk_field_symm_magsqr(long long n, const double *T, double *out) { launch(n, 256, fieldk::symm_magsqr_kernel, n, T, out); }

void hk_field_gather(int n, int nc, const int *cells, const double *f, double *out)
{
    launch((long long)n * nc, 256, fieldk::gather_kernel, n, nc, cells, f, out);
}

void hk_sngrad(int nFaces, int nc, const int *l, const int *u, const double *delta, const double *vf, double *out)
{
    launch((long long)nFaces * nc, 256, fieldk::sngrad_kernel, nFaces, nc, l, u, delta, vf, out);
}

void hk_limiter(int nFaces, int scheme, double twoByk, const int *l, const int *u, const double *faceFlux, const double *vf,
                const double *gradc, const double *C, double *out)
{
    launch(nFaces, 256, fieldk::limiter_kernel, nFaces, scheme, twoByk, l, u, faceFlux, vf, gradc, C, out);
}

void hk_limited_weights(long long n, const double *limiter, const double *cd, const double *faceFlux, double *out)
{
    launch(n, 256, fieldk::limited_weights_kernel, n, limiter, cd, faceFlux, out);
}

/* csrc/mules.cu: b200ldu_mules_limiter's three kernels in its launch order */
void hk_mules_limiter(const HostCase *h, int nIter, double rDeltaT, const double *rho, const double *rho0, const double *psi,
                      const double *psi0, const double *psiB, const double *phiBD, const double *phiBDB, const double *phiCorr,
                      const double *phiCorrB, const double *Sp, const double *Su, const double *V, double psiMax, double psiMin,
                      double *lambda, double *lambdaB, double *scratch /* 6*nCells */, int nCoupled, int corr, double extrema)
{
    using namespace mulesk;
    const int n = h->nCells, nF = h->nFaces, nB = h->nB;
    double *psiMaxn = scratch, *psiMinn = psiMaxn + n, *sumPhip = psiMinn + n, *mSumPhim = sumPhip + n, *lambdam = mSumPhim + n,
           *lambdap = lambdam + n;
    const int *bs = nB ? h->bStart : nullptr;
    launch(n, 128, mules_bounds_kernel, n, h->ownerStart, h->u, h->losortStart, h->losort, h->l, bs, h->bFaces, psi, psiB, phiBD,
           phiBDB, phiCorr, phiCorrB, psi0, rho, rho0, Sp, Su, V, rDeltaT, psiMax, psiMin, psiMaxn, psiMinn, sumPhip, mSumPhim, corr, extrema);
    for (int j = 0; j < nIter; j++) {
        launch(n, 128, mules_cell_lambda_kernel, n, h->ownerStart, h->losortStart, h->losort, bs, h->bFaces, lambda, lambdaB,
               phiCorr, phiCorrB, psiMaxn, psiMinn, sumPhip, mSumPhim, lambdam, lambdap);
        launch((long long)nF + nB, 256, mules_face_lambda_kernel, nF, nB, nCoupled, corr, h->l, h->u, h->bFaceCells, phiCorr, phiCorrB, phiBDB,
               lambdam, lambdap, lambda, lambdaB);
    }
}

/* csrc/fv.cu: the cell-parallel face sums (fv_kernels.cuh) */
void hk_surface_integrate(const HostCase *h, int nc, const double *ssf, const double *bssf, const double *V, double *out, int divideByV,
                          int neiSign)
{
    const int *bs = h->nB ? h->bStart : nullptr;
    if (nc == 1)
        launch(h->nCells, 128, fvk::surface_integrate_kernel<1>, h->nCells, h->ownerStart, h->losortStart, h->losort, bs, h->bFaces, ssf,
               bssf, V, out, divideByV, neiSign);
    else
        launch(h->nCells, 128, fvk::surface_integrate_kernel<3>, h->nCells, h->ownerStart, h->losortStart, h->losort, bs, h->bFaces, ssf,
               bssf, V, out, divideByV, neiSign);
}
void hk_gauss_grad(const HostCase *h, int nc, const double *Sf, const double *ssf, const double *bSf, const double *bssf, const double *V,
                   double *out)
{
    const int *bs = h->nB ? h->bStart : nullptr;
    if (nc == 1)
        launch(h->nCells, 128, fvk::gauss_grad_kernel<1>, h->nCells, h->ownerStart, h->losortStart, h->losort, bs, h->bFaces, Sf, ssf, bSf,
               bssf, V, out);
    else
        launch(h->nCells, 128, fvk::gauss_grad_kernel<3>, h->nCells, h->ownerStart, h->losortStart, h->losort, bs, h->bFaces, Sf, ssf, bSf,
               bssf, V, out);
}
void hk_grad_linear(const HostCase *h, int nc, const double *Sf, const double *w, const double *vf, const double *bSf, const double *bvf,
                    const double *V, double *out)
{
    const int *bs = h->nB ? h->bStart : nullptr;
    if (nc == 1)
        launch(h->nCells, 128, fvk::grad_linear_kernel<1>, h->nCells, h->ownerStart, h->u, h->losortStart, h->losort, h->l, bs, h->bFaces,
               Sf, w, vf, bSf, bvf, V, out);
    else
        launch(h->nCells, 128, fvk::grad_linear_kernel<3>, h->nCells, h->ownerStart, h->u, h->losortStart, h->losort, h->l, bs, h->bFaces,
               Sf, w, vf, bSf, bvf, V, out);
}
void hk_neg_sum_diag(const HostCase *h, const double *upper, const double *lower, double *diag)
{
    launch(h->nCells, 128, fvk::neg_sum_diag_kernel, h->nCells, h->ownerStart, h->losortStart, h->losort, upper, lower, diag);
}
}
