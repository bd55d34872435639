/*
 * harness.cpp -- runs the REFERENCE'S OWN lduMatrix code on the CPU.  TEST INFRASTRUCTURE ONLY.
 *
 * The translation unit includes, by path from /root/reference (nothing is copied into this repo):
 *   LDU/lduMatrix/lduMatrixATmul.C            lduMatrix::Amul, Tmul, sumA, residual, H1 and the
 *                                             matrixMultiplyFunctor they launch            (:42-554)
 *   LDU/lduMatrix/lduMatrixTemplates.C        lduMatrix::H, faceH                          (:50-160)
 *   LDU/lduAddressing/lduAddressingFunctors.H lduAddressingFunctor family, matrixOperation (:10-400)
 *   LDU/lduMatrix/lduMatrixFunctors.H         lduMatrixDiagonalResidualFunctor, ...        (:6-301)
 *   LDU/preconditioners/AINVPreconditioner/AINVPreconditionerF.H                           (:5-100)
 *   LDU/smoothers/Jacobi/JacobiSmootherF.H                                                 (:8-110)
 *   LDU/lduMatrix/lduMatrixSolverFunctors.H   the solvers' AXPY functors                   (:7-45)
 * against oracle/ref_harness/shim/ (host-memory stand-ins for gpuList, tmp, textures, lduAddressing,
 * lduMatrix's data members; thrust's host back end).  Built by `make -C oracle ref` with
 * -ffp-contract=off into oracle/_ref/libref_ldu.so, where /root/reference exists.
 *
 * What a bit-exact match with the oracle then means: the oracle's row arithmetic (order of the
 * products and sums) is the reference's, for the instructions a compiler emits WITHOUT fusing a*b+c.
 * nvcc's device build of the reference does fuse where an expression allows it; DESIGN.md section 2
 * lists where that applies (tails beyond three faces per side, the AXPY functors).
 */
#include "lduMatrix.H" /* the shim (first on the include path) */

#include <algorithm>

#include "lduMatrixATmul.C"     /* reference, through oracle/_ref/inc/ */
#include "lduMatrixTemplates.C" /* reference: H, faceH */

#include "AINVPreconditionerF.H"
#include "JacobiSmootherF.H"
#include "lduMatrixSolverFunctors.H"

int Foam::lduMatrixSolutionCache::favourSpeed = 0;

using namespace Foam;

namespace
{
struct Case {
    lduAddressing addr;
    lduMatrix m;
    scalargpuField lower, upper, diag, lowerSort, upperSort;
    std::vector<label> ownerSort;
    std::vector<scalar> ls, us;

    Case(int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart, const int *losort,
         const double *dg, const double *up, const double *lo)
        : lower(lo ? lo : up, nF), upper(up, nF), diag(dg, n), ownerSort(nF), ls(nF), us(nF)
    {
        for (int k = 0; k < nF; k++) { // lduAddressing.C:373-400 ownerSortAddr; lduMatrix.C:404-436 sorted coefficients
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
        lowerSort.view(ls.data(), nF);
        upperSort.view(us.data(), nF);
        m.addr_ = &addr;
        m.lowerPtr_ = lo ? &lower : nullptr; // symmetric matrices have no lower array (lduMatrix.C:328-345)
        m.upperPtr_ = &upper;
        m.diagPtr_ = &diag;
        m.lowerSortPtr_ = &lowerSort;
        m.upperSortPtr_ = &upperSort;
        m.level_ = 0;
        m.coarsest_ = false;
    }
};
} // namespace

#define CASE_ARGS int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart, \
                  const int *losort, const double *dg, const double *up, const double *lo
#define CASE_PASS n, nF, l, u, ownerStart, losortStart, losort, dg, up, lo

extern "C" {

/* op: 0 Amul, 1 Tmul, 2 sumA, 3 residual (x = psi, b = source), 4 H1, 5 H, 6 faceH (out has nF entries),
 * 7 negSumDiag, 8 sumDiag, 9 sumMagOffDiag (the compositions of lduMatrixOperations.C:36-104, which lives in a
 * file with the matrix-algebra operators and is not included whole).  favourSpeed selects the
 * reference's path: 0 = losort indirection, 1/2 = pre-sorted coefficients ("fast"). */
int ref_matrix_op(int op, int favourSpeed, CASE_ARGS, const double *x, const double *b, double *out)
{
    Case c(CASE_PASS);
    lduMatrixSolutionCache::favourSpeed = favourSpeed;
    scalargpuField o(out, n);
    FieldField<gpuField, scalar> noCoeffs(0);
    lduInterfaceFieldPtrsList noInterfaces;
    switch (op) {
    case 0: {
        scalargpuField psi(x, n);
        c.m.Amul(o, tmp<scalargpuField>(psi), noCoeffs, noInterfaces, 0);
        return 0;
    }
    case 1: {
        scalargpuField psi(x, n);
        c.m.Tmul(o, tmp<scalargpuField>(psi), noCoeffs, noInterfaces, 0);
        return 0;
    }
    case 2:
        c.m.sumA(o, noCoeffs, noInterfaces);
        return 0;
    case 3: {
        scalargpuField psi(x, n), src(b, n);
        c.m.residual(o, psi, src, noCoeffs, noInterfaces, 0);
        return 0;
    }
    case 4:
        c.m.H1(o);
        return 0;
    case 5: {
        scalargpuField psi(x, n);
        c.m.H(o, psi);
        return 0;
    }
    case 6: {
        scalargpuField psi(x, n), fo(out, nF);
        c.m.faceH(fo, psi);
        return 0;
    }
    case 7: // lduMatrixOperations.C:57-79 (diag starts from the values in `out`)
        matrixOperation(o.begin(), o, c.addr,
                        matrixCoeffsFunctor<scalar, negateUnaryOperatorFunctor<scalar, scalar>>(
                            c.m.lower().data(), negateUnaryOperatorFunctor<scalar, scalar>()),
                        matrixCoeffsFunctor<scalar, negateUnaryOperatorFunctor<scalar, scalar>>(
                            c.m.upper().data(), negateUnaryOperatorFunctor<scalar, scalar>()));
        return 0;
    case 8: // :36-55
        matrixOperation(o.begin(), o, c.addr,
                        matrixCoeffsFunctor<scalar, unityOp<scalar>>(c.m.lower().data(), unityOp<scalar>()),
                        matrixCoeffsFunctor<scalar, unityOp<scalar>>(c.m.upper().data(), unityOp<scalar>()));
        return 0;
    case 9: // :81-104
        matrixOperation(o.begin(), o, c.addr,
                        matrixCoeffsFunctor<scalar, magUnaryFunctionFunctor<scalar, scalar>>(
                            c.m.upper().data(), magUnaryFunctionFunctor<scalar, scalar>()),
                        matrixCoeffsFunctor<scalar, magUnaryFunctionFunctor<scalar, scalar>>(
                            c.m.lower().data(), magUnaryFunctionFunctor<scalar, scalar>()));
        return 0;
    }
    return -1;
}

/* coupledFvPatchField::updateInterfaceMatrix (coupledFvPatchField.C:236-257): result[faceCells] (-/+)= coeffs*pnf
 * through the reference's matrixPatchOperation + matrixInterfaceFunctor (lduAddressingFunctors.H:230-262,
 * 340-400).  The per-patch sort addressing (unique cells, faces of a cell in ascending patch-face order,
 * lduAddressing.C:38-130: stable sort by cell) is rebuilt here. */
int ref_interface_update(int nCells, int nPatchFaces, const int *faceCells, const double *coeffs, const double *pnf,
                         int negate, double *result_io)
{
    std::vector<label> order(nPatchFaces), cells, start;
    for (int i = 0; i < nPatchFaces; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](label a, label b) { return faceCells[a] < faceCells[b]; });
    for (int k = 0; k < nPatchFaces; k++) {
        const label c = faceCells[order[k]];
        if (cells.empty() || cells.back() != c) {
            cells.push_back(c);
            start.push_back(k);
        }
    }
    start.push_back(nPatchFaces);
    lduAddressing addr;
    addr.nCells_ = nCells;
    addr.patchCells_.view(cells.data(), (label)cells.size());
    addr.patchSort_.view(order.data(), nPatchFaces);
    addr.patchSortStart_.view(start.data(), (label)start.size());
    scalargpuField result(result_io, nCells);
    if (negate)
        matrixPatchOperation(0, result, addr, matrixInterfaceFunctor<scalar, true>(coeffs, pnf));
    else
        matrixPatchOperation(0, result, addr, matrixInterfaceFunctor<scalar, false>(coeffs, pnf));
    return 0;
}

/* AINVPreconditionerFunctor<fast,3> exactly as AINVPreconditioner.C:78-117 launches it: transpose swaps
 * the coefficient arrays (preconditionT, :119-160); rD = 1/diag (calcReciprocalD, :47-62). */
int ref_ainv(int fast, int transpose, CASE_ARGS, const double *r, double *w)
{
    Case c(CASE_PASS);
    std::vector<scalar> rD(n);
    for (int i = 0; i < n; i++) rD[i] = 1.0 / dg[i];
    const scalar *L = fast ? c.lowerSort.data() : c.lower.data();
    const scalar *U = c.upper.data();
    const scalar *Lt = c.lower.data(), *Ut = fast ? c.upperSort.data() : c.upper.data();
    const label *own = fast ? c.addr.ownerSort_.data() : c.addr.lower_.data();
    textures<scalar> rt(r), rDt(rD.data());
    if (fast) {
        AINVPreconditionerFunctor<true, 3> f(rt, rDt, transpose ? Ut : L, transpose ? Lt : U, own, c.addr.upper_.data(),
                                             ownerStart, losortStart, losort);
        for (label i = 0; i < n; i++) w[i] = f(i);
    } else {
        AINVPreconditionerFunctor<false, 3> f(rt, rDt, transpose ? Ut : L, transpose ? Lt : U, own, c.addr.upper_.data(),
                                              ownerStart, losortStart, losort);
        for (label i = 0; i < n; i++) w[i] = f(i);
    }
    return 0;
}

/* JacobiSmootherFunctor<fast,3>, one sweep without interfaces (JacobiSmoother.C:103-140) */
int ref_jacobi(int fast, double omega, CASE_ARGS, const double *psi, const double *b, double *out)
{
    Case c(CASE_PASS);
    textures<scalar> pt(psi);
    const scalar *L = fast ? c.lowerSort.data() : c.lower.data();
    const label *own = fast ? c.addr.ownerSort_.data() : c.addr.lower_.data();
    if (fast) {
        JacobiSmootherFunctor<true, 3> f(omega, pt, dg, b, L, c.upper.data(), own, c.addr.upper_.data(), ownerStart,
                                         losortStart, losort);
        for (label i = 0; i < n; i++) out[i] = f(i);
    } else {
        JacobiSmootherFunctor<false, 3> f(omega, pt, dg, b, L, c.upper.data(), own, c.addr.upper_.data(), ownerStart,
                                          losortStart, losort);
        for (label i = 0; i < n; i++) out[i] = f(i);
    }
    return 0;
}
}
