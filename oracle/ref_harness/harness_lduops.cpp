/* harness_lduops.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the reference's matrix-algebra operators
 *   LDU/lduMatrix/lduMatrixOperations.C   (operator=, negate, operator+=, operator-=, operator*=; sumDiag, negSumDiag, sumMagOffDiag)
 * for the host against oracle/ref_harness/shim/foam_shim.h (SHIM_REFERENCE_MATRIX_OPERATIONS: the declarations and the
 * allocate-on-demand accessors of lduMatrix.C:219-270), and exposes the combination the momentum equation of icoFoam forms:
 * a diagonal matrix (fvm::ddt) += an asymmetric one (fvm::div) -= a symmetric one (fvm::laplacian), any of the three kinds at
 * each place.  LDU/ = src/OpenFOAM/matrices/lduMatrix/.  Built by `make -C oracle ref` into oracle/_ref/libref_lduops.so. */
#define SHIM_REFERENCE_MATRIX_OPERATIONS
#include "lduMatrix.H" /* the shim */

#include <algorithm>
#include <stdexcept>
#include <vector>

namespace Foam
{
template <class A, class B, class R> struct multiplyOperatorFunctor { // gpuFieldM.H: a*b
    SHIM_HD R operator()(const A &a, const B &b) const { return a * b; }
};
struct WarningStream {
    template <class T> WarningStream &operator<<(const T &) { return *this; }
};
static WarningStream Warning;
#define WarningIn(where) ::Foam::Warning
static const char nl = '\n', endl = '\n';
} // namespace Foam

#include "lduMatrixOperations.C" /* reference, through oracle/_ref/inc_lduops/ */

int Foam::lduMatrixSolutionCache::favourSpeed = 0;
using namespace Foam;

namespace
{
struct Mat {
    lduMatrix m;
    static scalargpuField *owned(const double *p, int k) // a copy the operators may write to (the pointer constructor aliases)
    {
        scalargpuField *f = new scalargpuField(k, 0.0);
        std::copy(p, p + k, f->begin());
        return f;
    }
    Mat(const lduAddressing &a, int n, int nF, const double *d, const double *u, const double *l)
    {
        m.addr_ = &a;
        m.diagPtr_ = d ? owned(d, n) : nullptr;
        m.upperPtr_ = u ? owned(u, nF) : nullptr;
        m.lowerPtr_ = l ? owned(l, nF) : nullptr;
        m.lowerSortPtr_ = m.upperSortPtr_ = nullptr;
        m.level_ = 0;
        m.coarsest_ = false;
    }
};
} // namespace

extern "C" {
/* result = A (op1) B (op2) C with op = +1 (operator+=) or -1 (operator-=), 0 = skip; each matrix: diag [n] / upper [nF] /
 * lower [nF], NULL where the reference matrix has no such array.  has[3] <- which arrays the result holds. */
int ref_ldu_combine(int n, int nF, const int *l, const int *u, const double *dA, const double *uA, const double *lA, int op1,
                    const double *dB, const double *uB, const double *lB, int op2, const double *dC, const double *uC,
                    const double *lC, double *dOut, double *uOut, double *lOut, int *has)
{
    try {
        lduAddressing addr;
        addr.nCells_ = n;
        addr.lower_.view(l, nF);
        addr.upper_.view(u, nF);
        Mat A(addr, n, nF, dA, uA, lA), B(addr, n, nF, dB, uB, lB), C(addr, n, nF, dC, uC, lC);
        if (op1 > 0) A.m += B.m;
        if (op1 < 0) A.m -= B.m;
        if (op2 > 0) A.m += C.m;
        if (op2 < 0) A.m -= C.m;
        has[0] = A.m.diagPtr_ != nullptr;
        has[1] = A.m.upperPtr_ != nullptr;
        has[2] = A.m.lowerPtr_ != nullptr;
        if (has[0]) std::copy(A.m.diagPtr_->begin(), A.m.diagPtr_->end(), dOut);
        if (has[1]) std::copy(A.m.upperPtr_->begin(), A.m.upperPtr_->end(), uOut);
        if (has[2]) std::copy(A.m.lowerPtr_->begin(), A.m.lowerPtr_->end(), lOut);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
}
