/*
 * fvm_shim.h -- what the reference's fvMatrix.H / fvMatrix.C (included by path, never copied) need around them to
 * compile for the host: fields over plain memory, a GeometricField / fvPatchField / fvMesh reduced to what the member
 * functions exercised by harness_fvm.cpp touch, an lduMatrix with caller-filled coefficient arrays, dimensions and
 * streams as inert stand-ins.  TEST INFRASTRUCTURE ONLY.  Only the members instantiated by the harness are ever
 * compiled past the template definition; everything else of the 2000-line file just has to parse.
 */
#ifndef FVM_SHIM_H
#define FVM_SHIM_H
#include "fields_shim.h"

namespace Foam
{
// ---- dimensions: inert ----
struct dimensionSet {
    static int debug;
    dimensionSet() {}
    void operator+=(const dimensionSet &) {}
    void operator-=(const dimensionSet &) {}
    void operator*=(const dimensionSet &) {}
    void operator/=(const dimensionSet &) {}
    explicit dimensionSet(Istream &) {}
    bool operator==(const dimensionSet &) const { return true; }
    bool operator!=(const dimensionSet &) const { return false; }
    void reset(const dimensionSet &) {}
};
inline dimensionSet operator/(const dimensionSet &, const dimensionSet &) { return dimensionSet(); }
inline dimensionSet operator*(const dimensionSet &, const dimensionSet &) { return dimensionSet(); }
static const dimensionSet dimVol, dimless, dimVolume;
struct dimensionedScalarStub;
template <class T> struct dimensioned {
    word name_;
    T value_;
    dimensioned(const word &n, const dimensionSet &, const T &v) : name_(n), value_(v) {}
    dimensioned(const T &v) : value_(v) {}
    const word &name() const { return name_; }
    dimensionSet dimensions() const { return dimensionSet(); }
    const T &value() const { return value_; }
};
typedef dimensioned<scalar> dimensionedScalar;
struct IOobject {
    enum readOption { NO_READ };
    enum writeOption { NO_WRITE };
    template <class... A> IOobject(const A &...) {}
};
class dictionary
{
public:
    template <class T> bool readIfPresent(const word &, T &) const { return false; }
    template <class T> T lookupOrDefault(const word &, const T &d) const { return d; }
};
template <class T> class autoPtr
{
    mutable T *p_;

public:
    autoPtr(T *p = nullptr) : p_(p) {}
    autoPtr(const autoPtr &o) : p_(o.p_) { o.p_ = nullptr; }
    ~autoPtr() { delete p_; }
    T *operator->() const { return p_; }
    T &operator()() const { return *p_; }
};
struct Pstream {
    static bool master() { return true; }
};
struct UPstream {
    static int msgType() { return 0; }
};
template <class T, class Op> void reduce(T &, const Op &, int = 0, int = 0) {}
template <class T, class Op> T returnReduce(const T &v, const Op &, int = 0, int = 0) { return v; }
template <class T> struct sumOp {
};
template <class T> struct maxOp {
};

// ---- addressing and lduMatrix with caller-filled arrays ----
class lduAddressing
{
public:
    label nCells_;
    labelgpuList lower_, upper_, ownerStart_, losortStart_, losort_;
    // per-patch sort addressing (lduAddressing.C:38-167), built by the harness
    std::vector<labelgpuList> patchCells_, patchSort_, patchSortStart_;
    label size() const { return nCells_; }
    const labelgpuList &lowerAddr() const { return lower_; }
    const labelgpuList &upperAddr() const { return upper_; }
    const labelgpuList &ownerStartAddr() const { return ownerStart_; }
    const labelgpuList &losortStartAddr() const { return losortStart_; }
    const labelgpuList &losortAddr() const { return losort_; }
    const labelgpuList &patchSortCells(label p) const { return patchCells_[(size_t)p]; }
    const labelgpuList &patchSortAddr(label p) const { return patchSort_[(size_t)p]; }
    const labelgpuList &patchSortStartAddr(label p) const { return patchSortStart_[(size_t)p]; }
};
class fvMesh;
// the coupled patches as the component loop of solveSegregated sees them: result[faceCells] -= coeffs*pnf[cmpt]
// (coupledFvPatchField::updateInterfaceMatrix, lduAddressingFunctors.H:237-262 -- pinned through libref_ldu)
struct lduInterfaceFieldPtrsList {
    std::vector<std::function<void(const gpuField<scalar> &, gpuField<scalar> &, direction)>> update;
};
class solverPerformance
{
public:
    static int debug;
    solverPerformance() {}
    solverPerformance(const word &, const word &) {}
    word &solverName()
    {
        static word w;
        return w;
    }
    void print(Ostream &) const {}
};
inline solverPerformance max(const solverPerformance &a, const solverPerformance &) { return a; }

class lduMatrix
{
    const fvMesh &mesh_;
    scalargpuField diag_, upper_, lower_;
    bool hasLower_ = false;

public:
    // lduMatrix::solver::New(...)->solve(psi, source): the harness does not solve -- it records what the matrix and
    // the right-hand side look like at that moment (the folded diagonal, the total source) and leaves psi alone
    class solver
    {
        const lduMatrix &m_;

    public:
        static std::vector<scalar> seenDiag, seenSource;
        solver(const lduMatrix &m) : m_(m) {}
        virtual ~solver() {}
        static autoPtr<solver> New(const word &, const lduMatrix &m, const FieldField<gpuField, scalar> &,
                                   const FieldField<gpuField, scalar> &, const lduInterfaceFieldPtrsList &,
                                   const dictionary &)
        {
            return autoPtr<solver>(new solver(m));
        }
        void read(const dictionary &) {}
        solverPerformance solve(scalargpuField &, const scalargpuField &source, const direction = 0) const
        {
            seenDiag.insert(seenDiag.end(), m_.diag().begin(), m_.diag().end());      // one block per component solved
            seenSource.insert(seenSource.end(), source.begin(), source.end());
            return solverPerformance();
        }
    };
    lduMatrix(const fvMesh &mesh) : mesh_(mesh) {}
    lduMatrix(const lduMatrix &) = default;
    lduMatrix(lduMatrix &m, bool) : mesh_(m.mesh_), diag_(m.diag_), upper_(m.upper_), lower_(m.lower_), hasLower_(m.hasLower_) {}
    const lduAddressing &lduAddr() const;
    const fvMesh &mesh() const { return mesh_; }
    label level() const { return 0; }
    scalargpuField &diag() { return diag_; }
    const scalargpuField &diag() const { return diag_; }
    scalargpuField &upper() { return upper_; }
    const scalargpuField &upper() const { return upper_; }
    scalargpuField &lower()
    {
        hasLower_ = true;
        return lower_;
    }
    const scalargpuField &lower() const { return hasLower_ ? lower_ : upper_; }
    bool hasDiag() const { return diag_.size() > 0; }
    bool hasUpper() const { return upper_.size() > 0; }
    bool hasLower() const { return hasLower_; }
    bool symmetric() const { return hasUpper() && !hasLower_; }
    bool asymmetric() const { return hasLower_; }
    bool diagonal() const { return hasDiag() && !hasUpper() && !hasLower_; }
    void operator=(const lduMatrix &o)
    {
        diag_ = tmp<scalargpuField>(o.diag_);
        upper_ = tmp<scalargpuField>(o.upper_);
        lower_ = tmp<scalargpuField>(o.lower_);
        hasLower_ = o.hasLower_;
    }
    void negate()
    {
        diag_.negate();
        upper_.negate();
        lower_.negate();
    }
    // lduMatrixOperations.C:235-397, the kind analysis restated for this shim's storage (the reference's own operators are
    // pinned through libref_lduops; here they only carry fvMatrix::operator+= / -= of fvMatrix.C)
    void combine(const lduMatrix &A, const int sgn)
    {
        auto acc = [sgn](scalargpuField &t, const scalargpuField &a) {
            for (label i = 0; i < t.size(); i++) t.data()[i] = sgn > 0 ? t.data()[i] + a.data()[i] : t.data()[i] - a.data()[i];
        };
        auto take = [sgn](scalargpuField &t, const scalargpuField &a) {
            t = tmp<scalargpuField>(new scalargpuField(a.data(), a.size()));
            if (sgn < 0) t.negate();
        };
        if (A.hasDiag()) acc(diag_, A.diag_);
        if (symmetric() && A.symmetric())
            acc(upper_, A.upper_);
        else if (symmetric() && A.asymmetric()) {
            lower_ = tmp<scalargpuField>(new scalargpuField(upper_.data(), upper_.size()));
            hasLower_ = true;
            acc(upper_, A.upper_);
            acc(lower_, A.lower_);
        } else if (asymmetric() && A.symmetric()) {
            acc(lower_, A.upper_);
            acc(upper_, A.upper_);
        } else if (asymmetric() && A.asymmetric()) {
            acc(lower_, A.lower_);
            acc(upper_, A.upper_);
        } else if (diagonal()) {
            if (A.hasUpper()) take(upper_, A.upper_);
            if (A.hasLower_) {
                take(lower_, A.lower_);
                hasLower_ = true;
            }
        }
    }
    void operator+=(const lduMatrix &A) { combine(A, +1); }
    void operator-=(const lduMatrix &A) { combine(A, -1); }
    void operator*=(const scalargpuField &) { throw std::runtime_error("not used by the harness"); }
    void operator*=(scalar) { throw std::runtime_error("not used by the harness"); }
    // lduMatrixOperations.C:82-104 in the order its functors run (owner side: |upper|, neighbour side: |lower|)
    void sumMagOffDiag(scalargpuField &sumOff) const;
    // lduMatrixOperations.C:59-80 in the order its functors run (owner side: -lower, neighbour side: -upper)
    void negSumDiag();
    // lduMatrixTemplates.C:50-149 (pinned separately through libref_ldu)
    template <class Type> tmp<gpuField<Type>> H(const gpuField<Type> &psi) const;
    template <class Type> void H(gpuField<Type> &, const gpuField<Type> &) const;
    template <class Type> void faceH(gpuField<Type> &, const gpuField<Type> &) const;
    tmp<scalargpuField> H1() const { throw std::runtime_error("not used by the harness"); }
    void H1(scalargpuField &) const { throw std::runtime_error("not used by the harness"); }
    // lduMatrixATmul.C:397-496: rA = source - diag*psi - sum(off-diagonal*psi), then the coupled interfaces with the
    // negated coefficients of :455-463 (pinned through libref_ldu / libref_procfield; restated here for fvMatrix.C only)
    void residual(scalargpuField &rA, const scalargpuField &psi, const scalargpuField &source,
                  const FieldField<gpuField, scalar> &bouCoeffs, const lduInterfaceFieldPtrsList &ifs, const direction cmpt) const
    {
        const lduAddressing &a = lduAddr();
        for (label c = 0; c < psi.size(); c++) {
            scalar out = source.data()[c] - diag_.data()[c] * psi.data()[c];
            for (label f = a.ownerStart_.data()[c]; f < a.ownerStart_.data()[c + 1]; f++) {
                scalar p = upper().data()[f] * psi.data()[a.upper_.data()[f]];
                out = out + (-p);
            }
            for (label k = a.losortStart_.data()[c]; k < a.losortStart_.data()[c + 1]; k++) {
                const label f = a.losort_.data()[k];
                scalar p = lower().data()[f] * psi.data()[a.lower_.data()[f]];
                out = out + (-p);
            }
            rA.data()[c] = out;
        }
        for (size_t p = 0; p < ifs.update.size(); p++)
            if (ifs.update[p]) {
                gpuField<scalar> m(bouCoeffs[(label)p].data(), bouCoeffs[(label)p].size());
                m.negate();
                ifs.update[p](m, rA, cmpt);
            }
    }
    void initMatrixInterfaces(const FieldField<gpuField, scalar> &, const lduInterfaceFieldPtrsList &, const scalargpuField &,
                              scalargpuField &, const direction) const
    {
    }
    void updateMatrixInterfaces(const FieldField<gpuField, scalar> &coeffs, const lduInterfaceFieldPtrsList &ifs,
                                const scalargpuField &, scalargpuField &result, const direction cmpt) const
    {
        for (size_t p = 0; p < ifs.update.size(); p++)
            if (ifs.update[p]) ifs.update[p](coeffs[(label)p], result, cmpt);
    }
};

// ---- mesh, patches, fields ----
struct volMesh {
};
struct surfaceMesh {
};
template <class Type> class fvPatchField : public gpuField<Type>
{
public:
    const labelgpuList *faceCells_ = nullptr;
    const gpuField<Type> *internal_ = nullptr;
    bool coupled_ = false;
    gpuField<Type> pnf_;
    using gpuField<Type>::gpuField;
    using gpuField<Type>::operator=;
    bool coupled() const { return coupled_; }
    // boundary-condition coefficient functions (kind_: 0 fixedValue, fixedValueFvPatchField.C:113-146; 1 zeroGradient,
    // zeroGradientFvPatchField.C:111-150; coupled patches: coupledFvPatchField.C:162-209); patchDelta_ = patch().deltaCoeffs()
    int kind_ = 0;
    gpuField<scalar> patchDelta_;
    bool fixesValue() const { return !coupled_ && kind_ == 0; } // fixedValueFvPatchField.H fixesValue
    tmp<gpuField<Type>> valueInternalCoeffs(const tmp<gpuField<scalar>> &w) const
    {
        gpuField<Type> *r = new gpuField<Type>(this->size(), pTraits<Type>::zero);
        if (coupled_)
            for (label i = 0; i < r->size(); i++) r->data()[i] = Type(pTraits<Type>::one) * w().data()[i];
        else if (kind_ == 1)
            *r = Type(pTraits<Type>::one);
        return tmp<gpuField<Type>>(r);
    }
    tmp<gpuField<Type>> valueBoundaryCoeffs(const tmp<gpuField<scalar>> &w) const
    {
        gpuField<Type> *r = new gpuField<Type>(this->size(), pTraits<Type>::zero);
        if (coupled_)
            for (label i = 0; i < r->size(); i++) r->data()[i] = Type(pTraits<Type>::one) * (1.0 - w().data()[i]);
        else if (kind_ == 0)
            for (label i = 0; i < r->size(); i++) r->data()[i] = this->data()[i];
        return tmp<gpuField<Type>>(r);
    }
    tmp<gpuField<Type>> gradientInternalCoeffs(const gpuField<scalar> &deltaCoeffs) const // coupled
    {
        gpuField<Type> *r = new gpuField<Type>(this->size());
        for (label i = 0; i < r->size(); i++) r->data()[i] = -Type(pTraits<Type>::one) * deltaCoeffs.data()[i];
        return tmp<gpuField<Type>>(r);
    }
    tmp<gpuField<Type>> gradientBoundaryCoeffs(const gpuField<scalar> &deltaCoeffs) const // coupled: -gradientInternalCoeffs
    {
        gpuField<Type> *r = new gpuField<Type>(this->size());
        for (label i = 0; i < r->size(); i++) r->data()[i] = -(-Type(pTraits<Type>::one) * deltaCoeffs.data()[i]);
        return tmp<gpuField<Type>>(r);
    }
    tmp<gpuField<Type>> gradientInternalCoeffs() const
    {
        gpuField<Type> *r = new gpuField<Type>(this->size(), pTraits<Type>::zero);
        if (kind_ == 0)
            for (label i = 0; i < r->size(); i++) r->data()[i] = -Type(pTraits<Type>::one) * patchDelta_.data()[i];
        return tmp<gpuField<Type>>(r);
    }
    tmp<gpuField<Type>> gradientBoundaryCoeffs() const
    {
        gpuField<Type> *r = new gpuField<Type>(this->size(), pTraits<Type>::zero);
        if (kind_ == 0)
            for (label i = 0; i < r->size(); i++) r->data()[i] = patchDelta_.data()[i] * this->data()[i];
        return tmp<gpuField<Type>>(r);
    }
    tmp<gpuField<Type>> patchInternalField() const
    {
        gpuField<Type> *r = new gpuField<Type>(faceCells_->size());
        forAll((*faceCells_), i) r->data()[i] = internal_->data()[faceCells_->data()[i]];
        return tmp<gpuField<Type>>(r);
    }
    tmp<gpuField<Type>> patchNeighbourField() const { return tmp<gpuField<Type>>(new gpuField<Type>(pnf_.data(), pnf_.size())); }
};
template <class Type> class fvsPatchField : public gpuField<Type>
{
public:
    using gpuField<Type>::gpuField;
    using gpuField<Type>::operator=;
    void operator=(const tmp<gpuField<Type>> &t) { gpuField<Type>::operator=(t); }
};
template <class Type> struct zeroGradientFvPatchField {
    static const word typeName;
};
typedef zeroGradientFvPatchField<scalar> zeroGradientFvPatchScalarField;
template <class Type> struct calculatedFvPatchField {
    static const word typeName;
};
struct fvPatch {
    label size_;
    labelgpuList faceCells_;
    label size() const { return size_; }
    const labelgpuList &faceCells() const { return faceCells_; }
    template <class T> tmp<gpuField<T>> patchInternalField(const gpuField<T> &f) const
    {
        gpuField<T> *r = new gpuField<T>(faceCells_.size());
        forAll(faceCells_, i) r->data()[i] = f.data()[faceCells_.data()[i]];
        return tmp<gpuField<T>>(r);
    }
};
struct polyPatchStub {
    label start() const { return 0; }
    label size() const { return 0; }
};
struct polyBoundaryMeshStub {
    label size() const { return 0; }
    polyPatchStub operator[](label) const { return polyPatchStub(); }
    label whichPatch(label) const { return -1; }
};
struct fvBoundaryMesh {
    std::vector<fvPatch> p_;
    label size() const { return (label)p_.size(); }
    const fvPatch &operator[](label i) const { return p_[(size_t)i]; }
};
struct VolumeField {
    scalargpuField f_;
    const scalargpuField &getField() const { return f_; }
};

template <class Type> class fvsPatchField;
template <class Type, template <class> class PatchField, class GeoMesh> class GeometricField;
class fvMesh
{
public:
    // geometry fields the scheme sources name (only parsed unless the harness points them somewhere)
    const GeometricField<scalar, fvsPatchField, surfaceMesh> *magSf_ = nullptr, *deltaCoeffs_ = nullptr, *weights_ = nullptr;
    const GeometricField<scalar, fvsPatchField, surfaceMesh> &weights() const { return *weights_; }
    const GeometricField<vector, fvsPatchField, surfaceMesh> *Sf_ = nullptr;
    const GeometricField<scalar, fvsPatchField, surfaceMesh> &magSf() const { return *magSf_; }
    const GeometricField<scalar, fvsPatchField, surfaceMesh> &deltaCoeffs() const { return *deltaCoeffs_; }
    const GeometricField<vector, fvsPatchField, surfaceMesh> &Sf() const { return *Sf_; }
    lduAddressing addr_;
    fvBoundaryMesh boundary_;
    VolumeField V_;
    // Time and the volumes as the ddt schemes read them
    struct TimeStub {
        scalar deltaT_ = 1;
        scalar deltaTValue() const { return deltaT_; }
        dimensioned<scalar> deltaT() const { return dimensioned<scalar>(deltaT_); }
        word timeName() const { return word("0"); }
    };
    TimeStub time_;
    const TimeStub &time() const { return time_; }
    bool moving() const { return false; }
    struct VscHolder {
        const gpuField<scalar> *f;
        const VscHolder &operator()() const { return *this; }
        const gpuField<scalar> &getField() const { return *f; }
    };
    VscHolder Vsc() const { return VscHolder{&V_.f_}; }
    VscHolder Vsc0() const { return VscHolder{&V_.f_}; }
    const lduAddressing &lduAddr() const { return addr_; }
    const labelgpuList &owner() const { return addr_.lower_; }
    const labelgpuList &neighbour() const { return addr_.upper_; }
    polyBoundaryMeshStub boundaryMesh() const { return polyBoundaryMeshStub(); }
    label nInternalFaces() const { return addr_.lower_.size(); }
    const fvBoundaryMesh &boundary() const { return boundary_; }
    const VolumeField &V() const { return V_; }
    bool fluxRequired(const word &) const { return true; }
    int comm() const { return 0; }
    template <class T> void setSolverPerformance(const word &, const T &) const {}
    const dictionary &solverDict(const word &) const
    {
        static dictionary d;
        return d;
    }
    bool relaxEquation(const word &) const { return false; }
    scalar equationRelaxationFactor(const word &) const { return 1; }
    struct data {
        template <class T> T lookupOrDefault(const word &, const T &d) const { return d; }
    };
    Vector<label> solutionD() const { return Vector<label>(1, 1, 1); }
};
inline const lduAddressing &lduMatrix::lduAddr() const { return mesh_.lduAddr(); }
// The row operations below are the ones pinned through libref_ldu (lduMatrixATmul.C, lduMatrixTemplates.C,
// lduMatrixOperations.C compiled from the reference); here they only serve fvMatrix.C and follow the same order:
// owner faces ascending, then neighbour faces in losort order, products rounded separately.
inline void lduMatrix::sumMagOffDiag(scalargpuField &sumOff) const
{
    const lduAddressing &a = lduAddr();
    for (label c = 0; c < a.size(); c++) {
        scalar out = sumOff.data()[c];
        for (label f = a.ownerStart_.data()[c]; f < a.ownerStart_.data()[c + 1]; f++) out = out + std::fabs(upper().data()[f]);
        for (label k = a.losortStart_.data()[c]; k < a.losortStart_.data()[c + 1]; k++)
            out = out + std::fabs(lower().data()[a.losort_.data()[k]]);
        sumOff.data()[c] = out;
    }
}
inline void lduMatrix::negSumDiag()
{
    const lduAddressing &a = lduAddr();
    if ((label)diag_.size() != a.size()) diag_ = tmp<scalargpuField>(new scalargpuField(a.size(), 0.0));
    const lduMatrix &cm = *this; // the const lower() aliases upper() for a symmetric matrix (lduMatrix.C:328-345)
    for (label c = 0; c < a.size(); c++) {
        scalar out = diag_.data()[c];
        for (label f = a.ownerStart_.data()[c]; f < a.ownerStart_.data()[c + 1]; f++) out = out - cm.lower().data()[f];
        for (label k = a.losortStart_.data()[c]; k < a.losortStart_.data()[c + 1]; k++) out = out - cm.upper().data()[a.losort_.data()[k]];
        diag_.data()[c] = out;
    }
}
template <class Type> void lduMatrix::H(gpuField<Type> &Hpsi, const gpuField<Type> &psi) const
{
    const lduAddressing &a = lduAddr();
    for (label c = 0; c < a.size(); c++) {
        Type out = pTraits<Type>::zero;
        for (label f = a.ownerStart_.data()[c]; f < a.ownerStart_.data()[c + 1]; f++) {
            Type p = upper().data()[f] * psi.data()[a.upper_.data()[f]];
            out = out + (-p);
        }
        for (label k = a.losortStart_.data()[c]; k < a.losortStart_.data()[c + 1]; k++) {
            const label f = a.losort_.data()[k];
            Type p = lower().data()[f] * psi.data()[a.lower_.data()[f]];
            out = out + (-p);
        }
        Hpsi.data()[c] = out;
    }
}
template <class Type> tmp<gpuField<Type>> lduMatrix::H(const gpuField<Type> &psi) const
{
    gpuField<Type> *r = new gpuField<Type>(psi.size());
    H(*r, psi);
    return tmp<gpuField<Type>>(r);
}
template <class Type> void lduMatrix::faceH(gpuField<Type> &out, const gpuField<Type> &psi) const
{
    const lduAddressing &a = lduAddr();
    for (label f = 0; f < a.lower_.size(); f++) {
        Type p1 = upper().data()[f] * psi.data()[a.upper_.data()[f]];
        Type p2 = lower().data()[f] * psi.data()[a.lower_.data()[f]];
        out.data()[f] = p1 - p2;
    }
}

template <class Type, template <class> class PatchField, class GeoMesh> class GeometricField
{
public:
    class GeometricBoundaryField
    {
    public:
        std::vector<PatchField<Type>> p_;
        label size() const { return (label)p_.size(); }
        PatchField<Type> &operator[](label i) { return p_[(size_t)i]; }
        const PatchField<Type> &operator[](label i) const { return p_[(size_t)i]; }
        void updateCoeffs() {}
        // boundary-field algebra named by scheme variants that are parsed but never run here
        template <class O> GeometricBoundaryField operator*(const O &) const { throw std::runtime_error("not used"); }
        template <class O> GeometricBoundaryField operator-(const O &) const { throw std::runtime_error("not used"); }
        friend GeometricBoundaryField operator*(scalar, const GeometricBoundaryField &) { throw std::runtime_error("not used"); }
        lduInterfaceFieldPtrsList scalarInterfaces() const
        {
            lduInterfaceFieldPtrsList l;
            l.update.resize(p_.size());
            for (size_t p = 0; p < p_.size(); p++) {
                if (!p_[p].coupled()) continue;
                const PatchField<Type> *pf = &p_[p];
                l.update[p] = [pf](const gpuField<scalar> &c, gpuField<scalar> &result, direction d) {
                    forAll((*pf->faceCells_), i)
                    {
                        const scalar v = c.data()[i] * component(pf->pnf_.data()[i], d);
                        result.data()[pf->faceCells_->data()[i]] = result.data()[pf->faceCells_->data()[i]] + (-v);
                    }
                };
            }
            return l;
        }
        lduInterfaceFieldPtrsList interfaces() const { return lduInterfaceFieldPtrsList(); }
    };
    const fvMesh *mesh_ = nullptr;
    gpuField<Type> internal_;
    GeometricBoundaryField boundary_;
    label eventNo_ = 0;
    bool needReference_ = false;
    const GeometricField *old_ = nullptr; // the old-time level, when the harness provides one
    const GeometricField &oldTime() const { return old_ ? *old_ : *this; }
    GeometricField() {}
    template <class... A> GeometricField(const IOobject &, const fvMesh &m, const A &...) : mesh_(&m), internal_(m.lduAddr().size())
    {
        if constexpr (std::is_same<GeoMesh, surfaceMesh>::value) { // a face field: internal faces + one entry per patch
            internal_.setSize(m.nInternalFaces());
            boundary_.p_.resize((size_t)m.boundary().size());
        }
    }
    word type() const { return word("field"); }
    GeometricField(const IOobject &, const tmp<GeometricField> &t) : mesh_(t().mesh_), internal_(t().internal_), boundary_(t().boundary_) {}
    GeometricField(const tmp<GeometricField> &t) : mesh_(t().mesh_), internal_(t().internal_), boundary_(t().boundary_) {}
    static const GeometricField &null()
    {
        static GeometricField n;
        return n;
    }
    const word &name() const
    {
        static word w("psi");
        return w;
    }
    word instance() const { return word("0"); }
    const fvMesh &db() const { return *mesh_; }
    const fvMesh &mesh() const { return *mesh_; }
    dimensionSet dimensions() const { return dimensionSet(); }
    label size() const { return internal_.size(); }
    gpuField<Type> &internalField() { return internal_; }
    const gpuField<Type> &internalField() const { return internal_; }
    gpuField<Type> &getField() { return internal_; }
    const gpuField<Type> &getField() const { return internal_; }
    GeometricBoundaryField &boundaryField() { return boundary_; }
    const GeometricBoundaryField &boundaryField() const { return boundary_; }
    label &eventNo() { return eventNo_; }
    bool needReference() const { return needReference_; }
    void correctBoundaryConditions() {}
    word select(bool) const { return name(); }
    template <class F> void replace(direction, const F &) {}
    void operator+=(const GeometricField &) {}
    void operator-=(const GeometricField &) {}   // face-flux corrections: never present in the harness
    tmp<GeometricField> operator-() const { throw std::runtime_error("not used by the harness"); }
    void rename(const word &) {}
};
template <class Type, class GeoMesh> class DimensionedField
{
public:
    gpuField<Type> f_;
    const gpuField<Type> &field() const { return f_; }
    const gpuField<Type> &getField() const { return f_; }
    const fvMesh *mesh_ = nullptr;
    word name() const { return word("su"); }
    const fvMesh &mesh() const
    {
        if (!mesh_) throw std::runtime_error("not used by the harness");
        return *mesh_;
    }
    dimensionSet dimensions() const { return dimensionSet(); }
};
typedef fvPatchField<scalar> fvPatchScalarField;
typedef fvsPatchField<scalar> fvsPatchScalarField;
typedef GeometricField<vector, fvPatchField, volMesh> volVectorField;
typedef GeometricField<scalar, fvPatchField, volMesh> volScalarField;
typedef GeometricField<scalar, fvsPatchField, surfaceMesh> surfaceScalarField;
typedef DimensionedField<scalar, volMesh> volScalarFieldDimensioned;

struct fvMatrixCache { // fvMatrixCache.H: process-lifetime scratch vectors; here fresh storage per call
    static scalargpuField first(label, label n) { return scalargpuField(n); }
    static scalargpuField second(label, label n) { return scalargpuField(n); }
    static scalargpuField third(label, label n) { return scalargpuField(n); }
};
} // namespace Foam
#endif
