/*
 * fvm_shim.h -- what the reference's fvMatrix.H / fvMatrix.C (included by path, never copied) need around them to
 * compile for the host: fields over plain memory, a GeometricField / fvPatchField / fvMesh reduced to what the member
 * functions exercised by harness_fvm.cpp touch, an lduMatrix with caller-filled coefficient arrays, dimensions and
 * streams as inert stand-ins.  TEST INFRASTRUCTURE ONLY.  Only the members instantiated by the harness are ever
 * compiled past the template definition; everything else of the 2000-line file just has to parse.
 */
#ifndef FVM_SHIM_H
#define FVM_SHIM_H

#define __host__
#define __device__
#define __HOST____DEVICE__
#define NoRepository

#include <cmath>
#include <cstddef>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <thrust/functional.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/permutation_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/iterator/zip_iterator.h>
#include <thrust/copy.h>
#include <thrust/fill.h>
#include <thrust/transform.h>
#include <thrust/tuple.h>

#define forAll(list, i) for (Foam::label i = 0; i < (list).size(); i++)

namespace Foam
{
typedef double scalar;
typedef int label;
typedef unsigned char direction;
static const scalar SMALL = 1e-15, VSMALL = 1e-300, GREAT = 1e15;

class word : public std::string
{
public:
    word() {}
    word(const char *s) : std::string(s) {}
    word(const std::string &s) : std::string(s) {}
};
inline word operator+(const word &a, const char *b) { return word(static_cast<const std::string &>(a) + b); }
inline word operator+(const char *a, const word &b) { return word(a + static_cast<const std::string &>(b)); }
inline word operator+(const word &a, char c) { return word(static_cast<const std::string &>(a) + c); }
inline word operator+(const word &a, const word &b) { return word(static_cast<const std::string &>(a) + static_cast<const std::string &>(b)); }

struct Ostream {
    template <class T> Ostream &operator<<(const T &) { return *this; }
    Ostream &masterStream(int) { return *this; }
    bool good() const { return true; }
    void check(const char *) {}
};
struct Istream {
};
static Ostream Info, Pout, FatalError, Warning;
static const char endl = '\n', nl = '\n';
static Ostream FatalIOError;
#define FatalIOErrorIn(where, ios) ::Foam::FatalIOError
#define FatalErrorIn(where) ::Foam::FatalError
#define InfoIn(where) ::Foam::Info
#define WarningIn(where) ::Foam::Warning
inline int abort(Ostream &) { throw std::runtime_error("FatalError"); }
inline int exit(Ostream &) { throw std::runtime_error("FatalError"); }
#define ClassName(name)                             \
    static const char *typeName_() { return name; } \
    static const ::Foam::word typeName;             \
    static int debug
#define notImplemented(what) throw std::runtime_error("notImplemented")

template <class T> struct pTraits;
template <> struct pTraits<scalar> {
    static constexpr scalar zero = 0.0, one = 1.0;
    enum { nComponents = 1 };
    static const char *componentNames[];
};
inline scalar mag(scalar x) { return std::fabs(x); }
inline scalar component(scalar x, direction) { return x; }
inline scalar cmptMultiply(scalar a, scalar b) { return a * b; }
inline scalar cmptMax(scalar x) { return x; }
inline scalar cmptMin(scalar x) { return x; }
inline scalar cmptMag(scalar x) { return std::fabs(x); }
inline scalar cmptAv(scalar x) { return x; }
using std::max;
using std::min;

// Vector<Cmpt> reduced to what fvMatrix<vector> touches; component-wise operators as VectorSpaceI.H defines them
template <class Cmpt> class Vector
{
public:
    Cmpt v_[3];
    enum { nComponents = 3, rank = 1 };
    typedef Vector<label> labelType;
    Vector() {}
    Vector(Cmpt x, Cmpt y, Cmpt z) : v_{x, y, z} {}
    Cmpt &operator[](direction d) { return v_[d]; }
    const Cmpt &operator[](direction d) const { return v_[d]; }
    void operator+=(const Vector &o)
    {
        for (int i = 0; i < 3; i++) v_[i] += o.v_[i];
    }
    void operator-=(const Vector &o)
    {
        for (int i = 0; i < 3; i++) v_[i] -= o.v_[i];
    }
    void operator/=(scalar s)
    {
        for (int i = 0; i < 3; i++) v_[i] /= s;
    }
    void operator*=(scalar s)
    {
        for (int i = 0; i < 3; i++) v_[i] *= s;
    }
};
typedef Vector<scalar> vector;
template <class V, int r> struct powProduct;
template <> struct powProduct<Vector<label>, 1> {
    typedef Vector<label> type;
};
template <> struct pTraits<Vector<label>> {
    static const Vector<label> zero;
};
inline const Vector<label> pTraits<Vector<label>>::zero(0, 0, 0);
inline Vector<label> pow(const Vector<label> &v, const Vector<label> &) { return v; }   // solutionD: every direction solved
template <> struct pTraits<vector> {
    static const vector zero, one;
    enum { nComponents = 3 };
    static const char *componentNames[];
};
inline const vector pTraits<vector>::zero(0, 0, 0);
inline const vector pTraits<vector>::one(1, 1, 1);
inline vector operator+(const vector &a, const vector &b) { return vector(a.v_[0] + b.v_[0], a.v_[1] + b.v_[1], a.v_[2] + b.v_[2]); }
inline vector operator-(const vector &a, const vector &b) { return vector(a.v_[0] - b.v_[0], a.v_[1] - b.v_[1], a.v_[2] - b.v_[2]); }
inline vector operator-(const vector &a) { return vector(-a.v_[0], -a.v_[1], -a.v_[2]); }
inline vector operator*(scalar s, const vector &a) { return vector(s * a.v_[0], s * a.v_[1], s * a.v_[2]); }
inline vector operator*(const vector &a, scalar s) { return vector(a.v_[0] * s, a.v_[1] * s, a.v_[2] * s); }
inline vector operator/(const vector &a, scalar s) { return vector(a.v_[0] / s, a.v_[1] / s, a.v_[2] / s); }
inline scalar component(const vector &a, direction d) { return a.v_[d]; }
inline vector cmptMultiply(const vector &a, const vector &b) { return vector(a.v_[0] * b.v_[0], a.v_[1] * b.v_[1], a.v_[2] * b.v_[2]); }
inline vector cmptMag(const vector &a) { return vector(std::fabs(a.v_[0]), std::fabs(a.v_[1]), std::fabs(a.v_[2])); }
inline scalar cmptMax(const vector &a) { return std::max(std::max(a.v_[0], a.v_[1]), a.v_[2]); }   // VectorSpaceI.H:402-423
inline scalar cmptMin(const vector &a) { return std::min(std::min(a.v_[0], a.v_[1]), a.v_[2]); }
inline scalar cmptAv(const vector &a) { return ((a.v_[0] + a.v_[1]) + a.v_[2]) / 3; }              // :428-447

struct refCount {
};
struct zero {
};

// ---- fields over host memory ----
template <class T> class tmp;
template <class T> class gpuList
{
protected:
    std::vector<T> v_;

public:
    typedef T *iterator;
    typedef const T *const_iterator;
    gpuList() {}
    explicit gpuList(label n) : v_((size_t)n) {}
    gpuList(label n, const T &x) : v_((size_t)n, x) {}
    gpuList(const T *p, label n) : v_(p, p + n) {}
    gpuList(const gpuList &parent, label n) : v_(parent.v_.begin(), parent.v_.begin() + n) {}
    explicit gpuList(Istream &) { throw std::runtime_error("no streams in the harness"); }
    label size() const { return (label)v_.size(); }
    void setSize(label n) { v_.resize((size_t)n); }
    T *data() { return v_.data(); }
    const T *data() const { return v_.data(); }
    iterator begin() { return v_.data(); }
    iterator end() { return v_.data() + v_.size(); }
    const_iterator begin() const { return v_.data(); }
    const_iterator end() const { return v_.data() + v_.size(); }
    T get(label i) const { return v_[(size_t)i]; }
    void set(label i, const T &x) { v_[(size_t)i] = x; }
    void operator=(const T &x)
    {
        for (auto &e : v_) e = x;
    }
};
template <class T> class gpuField : public gpuList<T>
{
public:
    using gpuList<T>::gpuList;
    using gpuList<T>::operator=;
    gpuField() {}
    gpuField(const tmp<gpuField<T>> &t);
    void operator=(const tmp<gpuField<T>> &t);
    void negate()
    {
        for (auto &e : this->v_) e = -e;
    }
    tmp<gpuField<scalar>> component(direction) const;
    void replace(direction d, const gpuField<scalar> &f);
    void replace(direction d, const tmp<gpuField<scalar>> &f);
    void operator+=(const gpuField &o)
    {
        for (label i = 0; i < this->size(); i++) this->v_[(size_t)i] += o.data()[i];
    }
    void operator-=(const gpuField &o)
    {
        for (label i = 0; i < this->size(); i++) this->v_[(size_t)i] -= o.data()[i];
    }
    void operator+=(const tmp<gpuField> &o);
    void operator-=(const tmp<gpuField> &o);
    void operator*=(const gpuField<scalar> &o)
    {
        for (label i = 0; i < this->size(); i++) this->v_[(size_t)i] *= o.data()[i];
    }
    void operator/=(const gpuField<scalar> &o)
    {
        for (label i = 0; i < this->size(); i++) this->v_[(size_t)i] /= o.data()[i];
    }
    void operator*=(const scalar s)
    {
        for (auto &e : this->v_) e *= s;
    }
    void operator/=(const scalar s)
    {
        for (auto &e : this->v_) e /= s;
    }
};
typedef gpuField<scalar> scalargpuField;
typedef gpuList<label> labelgpuList;
typedef gpuList<label> labelUList;
template <class T> using UList = gpuList<T>;
template <class T> class UIndirectList;
template <class A, class B, class R> struct multiplyOperatorFunctor {
    R operator()(const A &a, const B &b) const { return a * b; }
};
// coupled-matrix solve and component bookkeeping: named by the parts of the sources that only have to parse
template <class Type, class DType, class LUType> class LduMatrix;
template <class Type> class SolverPerformance;


template <class T> class tmp
{
    mutable T *owned_;
    const T *ref_;

public:
    tmp(T *p = nullptr) : owned_(p), ref_(p) {}
    tmp(const T &r) : owned_(nullptr), ref_(&r) {}
    tmp(const tmp &o) : owned_(o.owned_), ref_(o.ref_) { o.owned_ = nullptr; }
    ~tmp() { delete owned_; }
    const T &operator()() const { return *ref_; }
    T &operator()() { return *const_cast<T *>(ref_); }
    bool isTmp() const { return owned_ != nullptr; }
    T *ptr() const
    {
        T *p = owned_;
        owned_ = nullptr;
        return p;
    }
    void clear() const
    {
        delete owned_;
        owned_ = nullptr;
    }
};
template <class T> gpuField<T>::gpuField(const tmp<gpuField<T>> &t) : gpuList<T>(t().data(), t().size()) {}
template <class T> void gpuField<T>::operator=(const tmp<gpuField<T>> &t) { this->v_.assign(t().begin(), t().end()); }
template <class T> void gpuField<T>::operator+=(const tmp<gpuField<T>> &t) { *this += t(); }
template <class T> void gpuField<T>::operator-=(const tmp<gpuField<T>> &t) { *this -= t(); }
template <class T> tmp<gpuField<scalar>> gpuField<T>::component(direction d) const
{
    gpuField<scalar> *r = new gpuField<scalar>(this->size());
    for (label i = 0; i < this->size(); i++) r->data()[i] = Foam::component(this->data()[i], d);
    return tmp<gpuField<scalar>>(r);
}
template <> inline void gpuField<scalar>::replace(direction, const gpuField<scalar> &f) { this->v_.assign(f.begin(), f.end()); }
template <> inline void gpuField<vector>::replace(direction d, const gpuField<scalar> &f)
{
    for (label i = 0; i < this->size(); i++) this->v_[(size_t)i].v_[d] = f.data()[i];
}
template <class T> void gpuField<T>::replace(direction d, const tmp<gpuField<scalar>> &f) { replace(d, f()); }
template <class T> void component(scalargpuField &out, const gpuField<T> &f, direction d) { out = f.component(d); }
inline tmp<gpuField<vector>> operator*(const tmp<scalargpuField> &a, const gpuField<vector> &b)
{
    gpuField<vector> *r = new gpuField<vector>(b.size());
    for (label i = 0; i < b.size(); i++) r->data()[i] = a().data()[i] * b.data()[i];
    return tmp<gpuField<vector>>(r);
}
inline tmp<scalargpuField> cmptAv(const gpuField<vector> &f)
{
    scalargpuField *r = new scalargpuField(f.size());
    for (label i = 0; i < f.size(); i++) r->data()[i] = cmptAv(f.data()[i]);
    return tmp<scalargpuField>(r);
}
// the field algebra relax() and D()/A() spell out (gpuFieldFunctions.C: one rounded operation per element)
#define SHIM_BINOP(op)                                                                              \
    inline tmp<scalargpuField> operator op(const scalargpuField &a, const scalargpuField &b)        \
    {                                                                                               \
        scalargpuField *r = new scalargpuField(a.size());                                           \
        for (label i = 0; i < a.size(); i++) r->data()[i] = a.data()[i] op b.data()[i];            \
        return tmp<scalargpuField>(r);                                                              \
    }                                                                                               \
    inline tmp<scalargpuField> operator op(const tmp<scalargpuField> &a, const scalargpuField &b) { return a() op b; } \
    inline tmp<scalargpuField> operator op(const scalargpuField &a, const tmp<scalargpuField> &b) { return a op b(); } \
    inline tmp<scalargpuField> operator op(const tmp<scalargpuField> &a, const tmp<scalargpuField> &b) { return a() op b(); }
SHIM_BINOP(+)
SHIM_BINOP(-)
SHIM_BINOP(*)
SHIM_BINOP(/)
#undef SHIM_BINOP
inline tmp<scalargpuField> operator-(const scalargpuField &a)
{
    scalargpuField *r = new scalargpuField(a.size());
    for (label i = 0; i < a.size(); i++) r->data()[i] = -a.data()[i];
    return tmp<scalargpuField>(r);
}
inline tmp<scalargpuField> operator*(scalar s, const scalargpuField &b)
{
    scalargpuField *r = new scalargpuField(b.size());
    for (label i = 0; i < b.size(); i++) r->data()[i] = s * b.data()[i];
    return tmp<scalargpuField>(r);
}
inline tmp<scalargpuField> cmptAv(const scalargpuField &f) { return tmp<scalargpuField>(new scalargpuField(f.data(), f.size())); }
inline tmp<scalargpuField> cmptMultiply(const scalargpuField &a, const scalargpuField &b) { return a * b; }
inline tmp<scalargpuField> cmptMultiply(const scalargpuField &a, const tmp<scalargpuField> &b) { return a * b(); }
inline label max(const gpuList<label> &l)
{
    label m = l.size() ? l.data()[0] : 0;
    forAll(l, i) m = l.data()[i] > m ? l.data()[i] : m;
    return m;
}

template <template <class> class Field, class T> class FieldField
{
    std::vector<std::shared_ptr<Field<T>>> v_;

public:
    FieldField() {}
    explicit FieldField(label n) : v_((size_t)n) {}
    FieldField(const FieldField &o)
    {
        for (auto &p : o.v_) v_.push_back(p ? std::make_shared<Field<T>>(*p) : nullptr);
    }
    label size() const { return (label)v_.size(); }
    Field<T> &operator[](label i) { return *v_[(size_t)i]; }
    const Field<T> &operator[](label i) const { return *v_[(size_t)i]; }
    void set(label i, Field<T> *p) { v_[(size_t)i].reset(p); }
    void set(label i, const tmp<Field<T>> &t) { v_[(size_t)i] = std::make_shared<Field<T>>(t()); }
    FieldField<Field, scalar> component(direction d) const
    {
        FieldField<Field, scalar> r(size());
        forAll((*this), i) r.set(i, (*this)[i].component(d));
        return r;
    }
    void negate()
    {
        for (auto &p : v_) p->negate();
    }
    void operator+=(const FieldField &o)
    {
        forAll(o, i) * v_[(size_t)i] += o[i];
    }
    void operator-=(const FieldField &o)
    {
        forAll(o, i) * v_[(size_t)i] -= o[i];
    }
};

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
    void operator+=(const lduMatrix &) { throw std::runtime_error("not used by the harness"); }
    void operator-=(const lduMatrix &) { throw std::runtime_error("not used by the harness"); }
    void operator*=(const scalargpuField &) { throw std::runtime_error("not used by the harness"); }
    void operator*=(scalar) { throw std::runtime_error("not used by the harness"); }
    // lduMatrixOperations.C:82-104 in the order its functors run (owner side: |upper|, neighbour side: |lower|)
    void sumMagOffDiag(scalargpuField &sumOff) const;
    void negSumDiag() { throw std::runtime_error("not used by the harness"); }
    // lduMatrixTemplates.C:50-149 (pinned separately through libref_ldu)
    template <class Type> tmp<gpuField<Type>> H(const gpuField<Type> &psi) const;
    template <class Type> void H(gpuField<Type> &, const gpuField<Type> &) const;
    template <class Type> void faceH(gpuField<Type> &, const gpuField<Type> &) const;
    tmp<scalargpuField> H1() const { throw std::runtime_error("not used by the harness"); }
    void H1(scalargpuField &) const { throw std::runtime_error("not used by the harness"); }
    // lduMatrixATmul.C:397-496 without interfaces: rA = source - diag*psi - sum(off-diagonal*psi)
    void residual(scalargpuField &rA, const scalargpuField &psi, const scalargpuField &source,
                  const FieldField<gpuField, scalar> &, const lduInterfaceFieldPtrsList &, const direction) const
    {
        gpuField<scalar> h(psi.size());
        H(h, psi);
        for (label c = 0; c < psi.size(); c++) {
            // same association as the reference functor: (source - diag*psi) + (-u*psi) + ...
            scalar out = source.data()[c] - diag_.data()[c] * psi.data()[c];
            rA.data()[c] = out;
        }
        const lduAddressing &a = lduAddr();
        for (label c = 0; c < psi.size(); c++) {
            scalar out = rA.data()[c];
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

class fvMesh
{
public:
    lduAddressing addr_;
    fvBoundaryMesh boundary_;
    VolumeField V_;
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
    GeometricField() {}
    template <class... A> GeometricField(const IOobject &, const fvMesh &m, const A &...) : mesh_(&m), internal_(m.lduAddr().size()) {}
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
    void rename(const word &) {}
};
template <class Type, class GeoMesh> class DimensionedField
{
public:
    gpuField<Type> f_;
    const gpuField<Type> &field() const { return f_; }
    const gpuField<Type> &getField() const { return f_; }
    const fvMesh &mesh() const
    {
        throw std::runtime_error("not used by the harness");
    }
    dimensionSet dimensions() const { return dimensionSet(); }
};
typedef fvPatchField<scalar> fvPatchScalarField;
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
