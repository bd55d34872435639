/*
 * harness_fvm.cpp -- runs the REFERENCE'S OWN fvMatrix<scalar> glue on the CPU.  TEST INFRASTRUCTURE ONLY.  Included by
 * path from /root/reference (src/finiteVolume/fvMatrices/):
 *   fvMatrix/fvMatrix.H, fvMatrix/fvMatrix.C     addToInternalField + functors, addBoundaryDiag, addCmptAvBoundaryDiag,
 *                                                addBoundarySource, setReference, relax, D, A, flux
 *   fvMatrix/fvMatrixSolve.C                     (parsed; the component loop needs a vector type and is not instantiated)
 *   fvScalarMatrix/fvScalarMatrix.H, .C          solveSegregated, residual, H for scalars
 *   FV/finiteVolume/convectionSchemes/gaussConvectionScheme/gaussConvectionScheme.C   fvmDiv (coefficient fill + patch coefficients)
 *   FV/finiteVolume/laplacianSchemes/gaussLaplacianScheme/gaussLaplacianScheme.C      fvmLaplacianUncorrected
 *   FV/interpolation/surfaceInterpolation/surfaceInterpolationScheme/surfaceInterpolationScheme.C  interpolate(vf)
 *   FV/finiteVolume/ddtSchemes/ddtScheme/ddtScheme.C, EulerDdtScheme/EulerDdtScheme.C  fvmDdt, fvcDdtPhiCorr, fvcDdtPhiCoeff
 * against oracle/ref_harness/shim_fvm/.  The linear solver behind solveSegregated is a recorder: it keeps the diagonal and
 * the source it is handed, which is what the folding has to get right.
 */
#define protected public
#define private public
#include "fvm_shim.h"

#include "fvMatrix.H" /* reference (pulls fvMatrix.C, fvMatrixSolve.C, fvScalarMatrix.H) */
#undef protected
#undef private
#include "fvScalarMatrix.C" /* reference */
#include "schemes_shim.h"
#include "gaussConvectionScheme.C" /* reference: fvmDiv :76-115 */
#include "gaussLaplacianScheme.C"  /* reference: fvmLaplacianUncorrected :46-89 */
#include "ddt_shim.h"
namespace Foam { int surfaceInterpolation::debug = 0; }
#include "surfaceInterpolationScheme.C" /* reference: interpolate(vf) :376-400 -> interpolate(vf, weights) :272-373 */
#include "ddtScheme.C"      /* reference: fvcDdtPhiCoeff :139-174 */
#include "EulerDdtScheme.C" /* reference: fvmDdt :331-361, fvcDdtPhiCorr :523-551 */

#include <algorithm>

namespace Foam
{
std::vector<scalar> lduMatrix::solver::seenDiag, lduMatrix::solver::seenSource;
int solverPerformance::debug = 0;
int dimensionSet::debug = 0;
template <> int fvMatrix<scalar>::debug = 0;
template <> int fvMatrix<vector>::debug = 0;
template <> const word zeroGradientFvPatchField<vector>::typeName("zeroGradient");
const char *pTraits<vector>::componentNames[] = {"x", "y", "z"};
template <> const word zeroGradientFvPatchField<scalar>::typeName("zeroGradient");
const char *pTraits<scalar>::componentNames[] = {""};
} // namespace Foam
using namespace Foam;

namespace
{
template <class Type> struct Case { // Type = scalar or vector; component arrays interleaved (3 doubles per vector)
    fvMesh mesh;
    GeometricField<Type, fvPatchField, volMesh> psi;
    std::unique_ptr<fvMatrix<Type>> M;
    static gpuField<Type> fld(const double *p, label n) { return gpuField<Type>(reinterpret_cast<const Type *>(p), n); }
    Case(int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart, const int *losort,
         int nP, const int *patchStart, const int *faceCells, const int *coupled, const double *pnf, const double *V,
         const double *psiv, const double *diag, const double *upper, const double *lower, const double *source,
         const double *ic, const double *bc)
    {
        const int nc = sizeof(Type) / sizeof(double);
        lduAddressing &a = mesh.addr_;
        a.nCells_ = n;
        a.lower_ = labelgpuList(l, nF);
        a.upper_ = labelgpuList(u, nF);
        a.ownerStart_ = labelgpuList(ownerStart, n + 1);
        a.losortStart_ = labelgpuList(losortStart, n + 1);
        a.losort_ = labelgpuList(losort, nF);
        mesh.V_.f_ = scalargpuField(V, n);
        psi.mesh_ = &mesh;
        psi.internal_ = fld(psiv, n);
        psi.boundary_.p_.resize((size_t)nP);
        for (int p = 0; p < nP; p++) {
            const int s = patchStart[p], np = patchStart[p + 1] - s;
            fvPatch fp;
            fp.size_ = np;
            fp.faceCells_ = labelgpuList(faceCells + s, np);
            mesh.boundary_.p_.push_back(fp);
            // per-patch sort addressing (lduAddressing.C:38-130): unique cells, faces sorted stably by cell
            std::vector<label> order(np), cells, start;
            for (int i = 0; i < np; i++) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](label x, label y) { return faceCells[s + x] < faceCells[s + y]; });
            for (int k = 0; k < np; k++) {
                const label c = faceCells[s + order[k]];
                if (cells.empty() || cells.back() != c) {
                    cells.push_back(c);
                    start.push_back(k);
                }
            }
            start.push_back(np);
            a.patchCells_.push_back(labelgpuList(cells.data(), (label)cells.size()));
            a.patchSort_.push_back(labelgpuList(order.data(), np));
            a.patchSortStart_.push_back(labelgpuList(start.data(), (label)start.size()));
        }
        for (int p = 0; p < nP; p++) { // after the vector stopped growing: the patch fields point into it
            const int s = patchStart[p], np = patchStart[p + 1] - s;
            fvPatchField<Type> &pf = psi.boundary_.p_[(size_t)p];
            pf.setSize(np);
            pf.faceCells_ = &mesh.boundary_.p_[(size_t)p].faceCells_;
            pf.internal_ = &psi.internal_;
            pf.coupled_ = coupled[p] != 0;
            pf.pnf_ = fld(pnf + (size_t)s * nc, np);
        }
        M.reset(new fvMatrix<Type>(psi, dimensionSet()));
        M->diag() = tmp<scalargpuField>(new scalargpuField(diag, n));
        M->upper() = tmp<scalargpuField>(new scalargpuField(upper, nF));
        if (lower) M->lower() = tmp<scalargpuField>(new scalargpuField(lower, nF));
        M->source() = tmp<gpuField<Type>>(new gpuField<Type>(fld(source, n)));
        for (int p = 0; p < nP; p++) {
            const int s = patchStart[p], np = patchStart[p + 1] - s;
            M->internalCoeffs()[p] = tmp<gpuField<Type>>(new gpuField<Type>(fld(ic + (size_t)s * nc, np)));
            M->boundaryCoeffs()[p] = tmp<gpuField<Type>>(new gpuField<Type>(fld(bc + (size_t)s * nc, np)));
        }
    }
};
void put(const scalargpuField &f, double *out) { std::copy(f.begin(), f.end(), out); }
void put(const gpuField<vector> &f, double *out)
{
    const double *p = reinterpret_cast<const double *>(f.data());
    std::copy(p, p + 3 * (size_t)f.size(), out);
}
} // namespace

/* surfaceInterpolationScheme<Type>::interpolate(vf) with the linear weights w (internal faces) / pw (patch faces): nc = 1 or 3;
 * patch values bvf, patchNeighbourField pnf (read on coupled patches) -> out [nF*nc], bout [tot*nc] */
template <class Type>
static int run_interpolate(int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart, const int *losort,
                           int nP, const int *patchStart, const int *faceCells, const int *coupled, const double *w, const double *pw,
                           const double *vf, const double *bvf, const double *pnf, double *out, double *bout)
{
    const int nc = sizeof(Type) / sizeof(double);
    const int tot = nP ? patchStart[nP] : 0;
    std::vector<double> zerosN((size_t)n * nc + 3, 0.0), zerosF((size_t)nF + 1, 0.0), zerosP((size_t)tot * nc + 3, 0.0), ones((size_t)n + 1, 1.0);
    Case<Type> C(n, nF, l, u, ownerStart, losortStart, losort, nP, patchStart, faceCells, coupled, pnf, ones.data(), vf, zerosN.data(),
                 zerosF.data(), nullptr, zerosN.data(), zerosP.data(), zerosP.data());
    for (int p = 0; p < nP; p++)
        static_cast<gpuField<Type> &>(C.psi.boundary_.p_[(size_t)p]) = tmp<gpuField<Type>>(
            new gpuField<Type>(reinterpret_cast<const Type *>(bvf + (size_t)patchStart[p] * nc), patchStart[p + 1] - patchStart[p]));
    surfaceScalarField weights;
    weights.mesh_ = &C.mesh;
    weights.internal_ = scalargpuField(w, nF);
    weights.boundary_.p_.resize((size_t)nP);
    for (int p = 0; p < nP; p++)
        static_cast<scalargpuField &>(weights.boundary_.p_[(size_t)p]) =
            tmp<scalargpuField>(new scalargpuField(pw + patchStart[p], patchStart[p + 1] - patchStart[p]));
    C.mesh.weights_ = &weights;
    tmp<GeometricField<Type, fvsPatchField, surfaceMesh>> t = surfaceInterpolationScheme<Type>(C.mesh).interpolate(C.psi);
    put(t().internal_, out);
    for (int p = 0; p < nP; p++) put(t().boundary_.p_[(size_t)p], bout + (size_t)patchStart[p] * nc);
    return 0;
}
extern "C" {
/* op: 0 addBoundaryDiag(x, 0) on x = in1 | 1 addCmptAvBoundaryDiag | 2 addBoundarySource(x, couples = iarg) | 3 setReference(cell
 * iarg, value darg, forced) -> out1 diag, out2 source | 4 relax(darg) -> out1 diag, out2 source | 5 D | 6 A | 7 flux -> out1
 * internal faces, out2 boundary faces (flat) | 8 H | 9 residual | 10 solveSegregated -> out1 diagonal and out2 source as the
 * solver sees them, out3 the diagonal afterwards */
int ref_fvm(int op, int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart, const int *losort,
            int nP, const int *patchStart, const int *faceCells, const int *coupled, const double *pnf, const double *V,
            const double *psi, const double *diag, const double *upper, const double *lower, const double *source,
            const double *ic, const double *bc, int iarg, double darg, const double *in1, double *out1, double *out2,
            double *out3)
{
    try {
        Case<scalar> C(n, nF, l, u, ownerStart, losortStart, losort, nP, patchStart, faceCells, coupled, pnf, V, psi, diag, upper, lower,
               source, ic, bc);
        fvMatrix<scalar> &M = *C.M;
        if (op == 0 || op == 1 || op == 2) {
            scalargpuField x(in1, n);
            if (op == 0) M.addBoundaryDiag(x, 0);
            if (op == 1) M.addCmptAvBoundaryDiag(x);
            if (op == 2) M.addBoundarySource(x, iarg != 0);
            put(x, out1);
        } else if (op == 3) {
            M.setReference(iarg, darg, true);
            put(M.diag(), out1);
            put(M.source(), out2);
        } else if (op == 4) {
            M.relax(darg);
            put(M.diag(), out1);
            put(M.source(), out2);
        } else if (op == 5) {
            put(M.D()(), out1);
        } else if (op == 6) {
            put(M.A()().internalField(), out1);
        } else if (op == 7) {
            GeometricField<scalar, fvsPatchField, surfaceMesh> phi;
            phi.mesh_ = &C.mesh;
            phi.internal_.setSize(nF);
            phi.boundary_.p_.resize((size_t)nP);
            for (int p = 0; p < nP; p++) phi.boundary_.p_[(size_t)p].setSize(patchStart[p + 1] - patchStart[p]);
            M.flux(phi);
            put(phi.internal_, out1);
            for (int p = 0; p < nP; p++) put(phi.boundary_.p_[(size_t)p], out2 + patchStart[p]);
        } else if (op == 8) {
            put(M.H()().internalField(), out1);
        } else if (op == 9) {
            put(M.residual()(), out1);
        } else if (op == 10) {
            lduMatrix::solver::seenDiag.clear();
            lduMatrix::solver::seenSource.clear();
            M.solveSegregated(dictionary());
            std::copy(lduMatrix::solver::seenDiag.begin(), lduMatrix::solver::seenDiag.end(), out1);
            std::copy(lduMatrix::solver::seenSource.begin(), lduMatrix::solver::seenSource.end(), out2);
            put(M.diag(), out3);
        } else
            return -2;
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

/* the same for fvMatrix<vector> (component arrays interleaved): op 0 addBoundaryDiag(x, cmpt = iarg) | 1 addCmptAvBoundaryDiag |
 * 2 addBoundarySource(x [n*3], couples = iarg) | 4 relax(darg) -> out1 diag, out2 source [n*3] | 5 D | 6 A | 8 H [n*3] (the
 * generic fvMatrix<Type>::H) | 10 solveSegregated -> out1 three diagonals, out2 three sources as the solver sees them, one block
 * per component, out3 the diagonal afterwards */
int ref_fvm_vec(int op, int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart,
                const int *losort, int nP, const int *patchStart, const int *faceCells, const int *coupled, const double *pnf,
                const double *V, const double *psi, const double *diag, const double *upper, const double *lower,
                const double *source, const double *ic, const double *bc, int iarg, double darg, const double *in1,
                double *out1, double *out2, double *out3)
{
    try {
        Case<vector> C(n, nF, l, u, ownerStart, losortStart, losort, nP, patchStart, faceCells, coupled, pnf, V, psi, diag, upper,
                       lower, source, ic, bc);
        fvMatrix<vector> &M = *C.M;
        if (op == 0 || op == 1) {
            scalargpuField x(in1, n);
            if (op == 0) M.addBoundaryDiag(x, (direction)iarg);
            if (op == 1) M.addCmptAvBoundaryDiag(x);
            put(x, out1);
        } else if (op == 2) {
            gpuField<vector> x(reinterpret_cast<const vector *>(in1), n);
            M.addBoundarySource(x, iarg != 0);
            put(x, out1);
        } else if (op == 4) {
            M.relax(darg);
            put(M.diag(), out1);
            put(M.source(), out2);
        } else if (op == 5) {
            put(M.D()(), out1);
        } else if (op == 6) {
            put(M.A()().internalField(), out1);
        } else if (op == 8) {
            put(M.H()().internalField(), out1);
        } else if (op == 10) {
            lduMatrix::solver::seenDiag.clear();
            lduMatrix::solver::seenSource.clear();
            M.solveSegregated(dictionary());
            std::copy(lduMatrix::solver::seenDiag.begin(), lduMatrix::solver::seenDiag.end(), out1);
            std::copy(lduMatrix::solver::seenSource.begin(), lduMatrix::solver::seenSource.end(), out2);
            put(M.diag(), out3);
        } else
            return -2;
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

/* Coefficient fills through the reference's schemes, for a scalar (nc 1) or vector (nc 3) field whose patches are fixedValue
 * (kind 0, values in pvalue), zeroGradient (kind 1) or coupled (kind 2):
 *   which 0: gaussConvectionScheme::fvmDiv(faceFlux, vf) with the given interpolation weights
 *   which 1: gaussLaplacianScheme::fvmLaplacianUncorrected(gammaMagSf, deltaCoeffs, vf)
 * a / b: internal-face fields (which 0: weights, faceFlux; which 1: gammaMagSf, deltaCoeffs); pa / pb: the same on the patch
 * faces (flat); pdelta: patch().deltaCoeffs() of the patch faces.  Outputs: lower (which 0 only), upper, diag, internalCoeffs,
 * boundaryCoeffs (flat, nc per face). */
int ref_fvm_fill(int which, int nc, int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart,
                 const int *losort, int nP, const int *patchStart, const int *faceCells, const int *kind, const double *pvalue,
                 const double *pdelta, const double *a, const double *b, const double *pa, const double *pb, double *lower,
                 double *upper, double *diag, double *ic, double *bc)
{
    try {
        std::vector<int> coupled((size_t)std::max(nP, 1));
        for (int p = 0; p < nP; p++) coupled[(size_t)p] = kind[p] == 2;
        const int tot = nP ? patchStart[nP] : 0;
        std::vector<double> zerosN((size_t)n * 3 + 3, 0.0), zerosF((size_t)nF + 1, 0.0), zerosP((size_t)tot * 3 + 3, 0.0), ones((size_t)n, 1.0);
        auto surf = [&](surfaceScalarField &f, const fvMesh &mesh, const double *in, const double *pin) {
            f.mesh_ = &mesh;
            f.internal_ = scalargpuField(in, nF);
            f.boundary_.p_.resize((size_t)nP);
            for (int p = 0; p < nP; p++)
                static_cast<scalargpuField &>(f.boundary_.p_[(size_t)p]) = tmp<scalargpuField>(
                    new scalargpuField(pin + patchStart[p], patchStart[p + 1] - patchStart[p]));
        };
        auto run = [&](auto &C, auto tag) {
            typedef decltype(tag) Type;
            const int k = sizeof(Type) / sizeof(double);
            for (int p = 0; p < nP; p++) {
                fvPatchField<Type> &pf = C.psi.boundary_.p_[(size_t)p];
                const int s = patchStart[p], np = patchStart[p + 1] - s;
                pf.kind_ = kind[p];
                pf.patchDelta_ = scalargpuField(pdelta + s, np);
                static_cast<gpuField<Type> &>(pf) = tmp<gpuField<Type>>(
                    new gpuField<Type>(reinterpret_cast<const Type *>(pvalue + (size_t)s * k), np));
            }
            surfaceScalarField A, B;
            surf(A, C.mesh, a, pa);
            surf(B, C.mesh, b, pb);
            tmp<fvMatrix<Type>> tM = which == 0
                                         ? fv::gaussConvectionScheme<Type>(C.mesh, B, tmp<surfaceInterpolationScheme<Type>>(
                                                                                          new surfaceInterpolationScheme<Type>(A)))
                                               .fvmDiv(B, C.psi)
                                         : fv::gaussLaplacianScheme<Type, scalar>::fvmLaplacianUncorrected(A, B, C.psi);
            fvMatrix<Type> &M = tM();
            if (which == 0) put(M.lower(), lower);
            put(M.upper(), upper);
            put(M.diag(), diag);
            for (int p = 0; p < nP; p++) {
                put(M.internalCoeffs()[p], ic + (size_t)patchStart[p] * k);
                put(M.boundaryCoeffs()[p], bc + (size_t)patchStart[p] * k);
            }
        };
        if (nc == 1) {
            Case<scalar> C(n, nF, l, u, ownerStart, losortStart, losort, nP, patchStart, faceCells, coupled.data(), zerosP.data(),
                           ones.data(), zerosN.data(), zerosN.data(), zerosF.data(), nullptr, zerosN.data(), zerosP.data(),
                           zerosP.data());
            run(C, scalar());
        } else {
            Case<vector> C(n, nF, l, u, ownerStart, losortStart, losort, nP, patchStart, faceCells, coupled.data(), zerosP.data(),
                           ones.data(), zerosN.data(), zerosN.data(), zerosF.data(), nullptr, zerosN.data(), zerosP.data(),
                           zerosP.data());
            run(C, vector());
        }
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

/* Euler ddt of a vector field U (old-time values U0 [n*3], boundary values per patch face [tot*3], fixesValue per patch):
 * EulerDdtScheme<vector>::fvmDdt(U) -> diag [n], source [n*3];  fvcDdtPhiCorr(U, phi) with phi.oldTime() = phi0 (internal
 * faces) / bphi0 (patch faces), face areas Sf / bSf, linear weights w -> ddtCorr on the internal faces [nF] and patch faces
 * [tot] */
int ref_ddt(int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart, const int *losort, int nP,
            const int *patchStart, const int *faceCells, const int *fixesValue, double deltaT, const double *V, const double *U0,
            const double *bU0, const double *phi0, const double *bphi0, const double *Sf, const double *bSf, const double *w,
            double *diag, double *source, double *ddtCorr, double *bddtCorr)
{
    try {
        const int tot = nP ? patchStart[nP] : 0;
        std::vector<int> coupled((size_t)std::max(nP, 1), 0);
        std::vector<double> zerosN((size_t)n * 3 + 3, 0.0), zerosF((size_t)nF + 1, 0.0), zerosP((size_t)tot * 3 + 3, 0.0);
        Case<vector> C(n, nF, l, u, ownerStart, losortStart, losort, nP, patchStart, faceCells, coupled.data(), zerosP.data(), V, U0,
                       zerosN.data(), zerosF.data(), nullptr, zerosN.data(), zerosP.data(), zerosP.data());
        C.mesh.time_.deltaT_ = deltaT;
        for (int p = 0; p < nP; p++) {
            fvPatchField<vector> &pf = C.psi.boundary_.p_[(size_t)p];
            const int s = patchStart[p], np = patchStart[p + 1] - s;
            pf.kind_ = fixesValue[p] ? 0 : 1;
            static_cast<gpuField<vector> &>(pf) = tmp<gpuField<vector>>(
                new gpuField<vector>(reinterpret_cast<const vector *>(bU0 + (size_t)s * 3), np));
        }
        auto surfS = [&](surfaceScalarField &f, const double *in, const double *pin) {
            f.mesh_ = &C.mesh;
            f.internal_ = scalargpuField(in, nF);
            f.boundary_.p_.resize((size_t)nP);
            for (int p = 0; p < nP; p++)
                static_cast<scalargpuField &>(f.boundary_.p_[(size_t)p]) =
                    tmp<scalargpuField>(new scalargpuField(pin + patchStart[p], patchStart[p + 1] - patchStart[p]));
        };
        surfaceScalarField phi, weights;
        surfS(phi, phi0, bphi0);
        std::vector<double> half((size_t)tot + 1, 0.5);
        surfS(weights, w, half.data());
        surfaceVectorField area;
        area.mesh_ = &C.mesh;
        area.internal_ = gpuField<vector>(reinterpret_cast<const vector *>(Sf), nF);
        area.boundary_.p_.resize((size_t)nP);
        for (int p = 0; p < nP; p++)
            static_cast<gpuField<vector> &>(area.boundary_.p_[(size_t)p]) = tmp<gpuField<vector>>(
                new gpuField<vector>(reinterpret_cast<const vector *>(bSf + (size_t)patchStart[p] * 3), patchStart[p + 1] - patchStart[p]));
        C.mesh.Sf_ = &area;
        C.mesh.weights_ = &weights;
        fv::EulerDdtScheme<vector> scheme(C.mesh);
        tmp<fvMatrix<vector>> tM = scheme.fvmDdt(C.psi);
        put(tM().diag(), diag);
        put(tM().source(), source);
        tmp<surfaceScalarField> tc = scheme.fvcDdtPhiCorr(C.psi, phi);
        put(tc().internal_, ddtCorr);
        for (int p = 0; p < nP; p++) put(tc().boundary_.p_[(size_t)p], bddtCorr + patchStart[p]);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

int ref_interpolate(int nc, int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart, const int *losort,
                    int nP, const int *patchStart, const int *faceCells, const int *coupled, const double *w, const double *pw,
                    const double *vf, const double *bvf, const double *pnf, double *out, double *bout)
{
    try {
        return nc == 1 ? run_interpolate<scalar>(n, nF, l, u, ownerStart, losortStart, losort, nP, patchStart, faceCells, coupled, w, pw,
                                                 vf, bvf, pnf, out, bout)
                       : run_interpolate<vector>(n, nF, l, u, ownerStart, losortStart, losort, nP, patchStart, faceCells, coupled, w, pw,
                                                 vf, bvf, pnf, out, bout);
    } catch (const std::exception &) {
        return -1;
    }
}

/* UEqn = (A + B - C) == su with the reference's fvMatrix operators (fvMatrix.C operator+ :1839, operator- :1990, operator==
 * :2252-2276, operator+= / -= :1750-1815) on three vector matrices over the same field: A diagonal (diag, source), B asymmetric
 * (diag, upper, lower, ic, bc), C symmetric (diag, upper, ic, bc); su [n*3].  Outputs: the coefficient arrays of the result. */
int ref_fvm_assemble(int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart, const int *losort, int nP,
                     const int *patchStart, const int *faceCells, const double *V, const double *aDiag, const double *aSource,
                     const double *bDiag, const double *bUpper, const double *bLower, const double *bIc, const double *bBc,
                     const double *cDiag, const double *cUpper, const double *cIc, const double *cBc, const double *su, double *diag,
                     double *upper, double *lower, double *source, double *ic, double *bc)
{
    try {
        const int tot = nP ? patchStart[nP] : 0;
        std::vector<int> coupled((size_t)std::max(nP, 1), 0);
        std::vector<double> zN((size_t)n * 3 + 3, 0.0), zF((size_t)nF + 1, 0.0), zP((size_t)tot * 3 + 3, 0.0);
        Case<vector> C(n, nF, l, u, ownerStart, losortStart, losort, nP, patchStart, faceCells, coupled.data(), zP.data(), V, zN.data(),
                       zN.data(), zF.data(), nullptr, zN.data(), zP.data(), zP.data());
        auto fld = [](const double *p, label k) { return tmp<gpuField<vector>>(new gpuField<vector>(reinterpret_cast<const vector *>(p), k)); };
        auto make = [&](const double *d, const double *up, const double *lo, const double *src, const double *i, const double *b) {
            tmp<fvMatrix<vector>> t(new fvMatrix<vector>(C.psi, dimensionSet()));
            fvMatrix<vector> &M = t();
            M.diag() = tmp<scalargpuField>(new scalargpuField(d, n));
            if (up) M.upper() = tmp<scalargpuField>(new scalargpuField(up, nF));
            if (lo) M.lower() = tmp<scalargpuField>(new scalargpuField(lo, nF));
            M.source() = fld(src ? src : zN.data(), n);
            for (int p = 0; p < nP; p++) {
                const int s = patchStart[p], np = patchStart[p + 1] - s;
                M.internalCoeffs()[p] = fld((i ? i : zP.data()) + (size_t)s * 3, np);
                M.boundaryCoeffs()[p] = fld((b ? b : zP.data()) + (size_t)s * 3, np);
            }
            return t;
        };
        tmp<fvMatrix<vector>> tA = make(aDiag, nullptr, nullptr, aSource, nullptr, nullptr);
        tmp<fvMatrix<vector>> tB = make(bDiag, bUpper, bLower, nullptr, bIc, bBc);
        tmp<fvMatrix<vector>> tC = make(cDiag, cUpper, nullptr, nullptr, cIc, cBc);
        DimensionedField<vector, volMesh> suF;
        suF.f_ = gpuField<vector>(reinterpret_cast<const vector *>(su), n);
        suF.mesh_ = &C.mesh;
        tmp<fvMatrix<vector>> tR = ((tA + tB) - tC) == suF;
        fvMatrix<vector> &R = tR();
        put(R.diag(), diag);
        put(R.upper(), upper);
        put(static_cast<const lduMatrix &>(R).lower(), lower);
        put(R.source(), source);
        for (int p = 0; p < nP; p++) {
            put(R.internalCoeffs()[p], ic + (size_t)patchStart[p] * 3);
            put(R.boundaryCoeffs()[p], bc + (size_t)patchStart[p] * 3);
        }
        return R.asymmetric() ? 2 : (R.symmetric() ? 1 : 0);
    } catch (const std::exception &) {
        return -1;
    }
}
}
