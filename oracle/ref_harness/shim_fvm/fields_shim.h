/*
 * fields_shim.h -- host stand-ins shared by the finite-volume harnesses: word, inert streams, scalar / Vector<Cmpt> with
 * the component-wise operators of VectorSpaceI.H, gpuList / gpuField / tmp / FieldField over plain memory with one rounded
 * operation per element and operator.  TEST INFRASTRUCTURE ONLY.
 */
#ifndef FIELDS_SHIM_H
#define FIELDS_SHIM_H

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
    bool eof() const { return true; }
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
inline tmp<gpuField<vector>> operator*(const scalargpuField &a, const gpuField<vector> &b) // V*su of fvMatrix.C operator==
{
    gpuField<vector> *r = new gpuField<vector>(b.size());
    for (label i = 0; i < b.size(); i++) r->data()[i] = a.data()[i] * b.data()[i];
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

} // namespace Foam
#endif
