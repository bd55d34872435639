/*
 * foam_shim.h -- the few OpenFOAM/RapidCFD types the reference's lduMatrix functor code is written
 * against, re-declared over plain host memory so that the REFERENCE'S OWN SOURCE FILES (included by
 * path from /root/reference, never copied) compile with g++ and run on the CPU.  TEST INFRASTRUCTURE
 * ONLY (oracle/Makefile target `ref`): it lets tests compare the oracle's restatement with the code it
 * restates.  Nothing here is an algorithm of the hot path; every loop that is executed comes from the
 * reference file named in harness.cpp.
 *
 * Device qualifiers are defined away and thrust runs its host (CPP) back end, so `thrust::transform`
 * over these containers is a serial loop calling the reference's functor once per cell.
 */
#ifndef FOAM_SHIM_H
#define FOAM_SHIM_H

#ifndef __CUDACC__
#define __host__
#define __device__
#define __HOST____DEVICE__
#define SHIM_HD
#else /* device probe (ref_harness/device_probe.cu): keep the qualifiers so that nvcc emits the functors */
#define __HOST____DEVICE__ __host__ __device__
#define SHIM_HD __host__ __device__
#endif

#include <cstddef>
#include <functional>
#include <stdexcept>
#include <vector>

#include <thrust/functional.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/permutation_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/iterator/zip_iterator.h>
#include <thrust/transform.h>
#include <thrust/tuple.h>

#define forAll(list, i) for (Foam::label i = 0; i < (list).size(); i++)

namespace Foam
{
typedef double scalar;       // etc/bashrc:76 WM_PRECISION_OPTION=DP
typedef int label;           // label.H:46-67 (32-bit)
typedef unsigned char direction;

// ---- gpuList / gpuField: a non-owning or owning view of host memory ----
template <class T> class gpuList
{
    std::vector<T> own_;
    T *p_;
    label n_;

public:
    typedef T *iterator;
    typedef const T *const_iterator;
    gpuList() : p_(nullptr), n_(0) {}
    explicit gpuList(label n) : own_((size_t)n), p_(own_.data()), n_(n) {}
    gpuList(label n, const T &v) : own_((size_t)n, v), p_(own_.data()), n_(n) {}
    gpuList(T *p, label n) : p_(p), n_(n) {}
    gpuList(const T *p, label n) : p_(const_cast<T *>(p)), n_(n) {}
    gpuList(const gpuList &o) : own_(o.p_, o.p_ + o.n_), p_(own_.data()), n_(o.n_) {} // deep copy
    gpuList(const gpuList &parent, label n) : p_(parent.p_), n_(n) {}                  // "delegate": a view
    gpuList &operator=(const gpuList &o)
    {
        if (n_ != o.n_ && (label)own_.size() == n_) setSize(o.n_); // an owning (or empty) list takes the size; a view does not
        for (label i = 0; i < n_ && i < o.n_; i++) p_[i] = o.p_[i];
        return *this;
    }
    void operator=(const T &v)
    {
        for (label i = 0; i < n_; i++) p_[i] = v;
    }
    void view(const T *p, label n) // shim only: point this list at caller memory
    {
        own_.clear();
        p_ = const_cast<T *>(p);
        n_ = n;
    }
    label size() const { return n_; }
    label byteSize() const { return n_ * (label)sizeof(T); }
    void setSize(label n)
    {
        own_.assign((size_t)n, T());
        p_ = own_.data();
        n_ = n;
    }
    void operator-=(const gpuList &o)
    {
        for (label i = 0; i < n_; i++) p_[i] -= o.p_[i];
    }
    void operator+=(const gpuList &o)
    {
        for (label i = 0; i < n_; i++) p_[i] += o.p_[i];
    }
    void operator*=(const gpuList &o)
    {
        for (label i = 0; i < n_; i++) p_[i] *= o.p_[i];
    }
    void operator*=(const T &s)
    {
        for (label i = 0; i < n_; i++) p_[i] *= s;
    }
    void negate()
    {
        for (label i = 0; i < n_; i++) p_[i] = -p_[i];
    }
    T *data() { return p_; }
    const T *data() const { return p_; }
    iterator begin() { return p_; }
    iterator end() { return p_ + n_; }
    const_iterator begin() const { return p_; }
    const_iterator end() const { return p_ + n_; }
};

template <class T> class tmp;
template <class T> class gpuField : public gpuList<T>
{
public:
    using gpuList<T>::gpuList;
    using gpuList<T>::operator=;
    gpuField() {}
    gpuField(const gpuField &o) : gpuList<T>(o) {}
    gpuField &operator=(const gpuField &o)
    {
        gpuList<T>::operator=(o);
        return *this;
    }
    gpuField(const gpuList<T> &f, const gpuList<label> &map) : gpuList<T>(map.size()) // gpuField(mapF, mapAddressing)
    {
        for (label i = 0; i < map.size(); i++) this->data()[i] = f.data()[map.data()[i]];
    }
    gpuField(const tmp<gpuField<T>> &t); // deep copy of a temporary (defined after tmp)
    void operator=(const tmp<gpuField<T>> &t);
};
typedef gpuField<scalar> scalargpuField;
typedef gpuList<label> labelgpuList;

// ---- tmp<T>: reference or owned temporary ----
template <class T> class tmp
{
    mutable T *owned_;
    const T *ref_;

public:
    tmp(T *p) : owned_(p), ref_(p) {}
    tmp(const T &r) : owned_(nullptr), ref_(&r) {}
    tmp(const tmp &o) : owned_(o.owned_), ref_(o.ref_) { o.owned_ = nullptr; }
    ~tmp() { delete owned_; }
    const T &operator()() const { return *ref_; }
    T &operator()() { return *const_cast<T *>(ref_); }
    void clear() const
    {
        delete owned_;
        owned_ = nullptr;
    }
};

template <class T> gpuField<T>::gpuField(const tmp<gpuField<T>> &t) : gpuList<T>(static_cast<const gpuList<T> &>(t())) {}
template <class T> void gpuField<T>::operator=(const tmp<gpuField<T>> &t)
{
    gpuList<T>::operator=(static_cast<const gpuList<T> &>(t()));
}

template <class T> tmp<gpuField<T>> operator-(const gpuField<T> &a, const gpuField<T> &b)
{
    gpuField<T> *r = new gpuField<T>(a.size());
    for (label i = 0; i < a.size(); i++) r->data()[i] = a.data()[i] - b.data()[i];
    return tmp<gpuField<T>>(r);
}
template <class T> tmp<gpuField<T>> operator/(const gpuField<T> &a, const gpuField<T> &b)
{
    gpuField<T> *r = new gpuField<T>(a.size());
    for (label i = 0; i < a.size(); i++) r->data()[i] = a.data()[i] / b.data()[i];
    return tmp<gpuField<T>>(r);
}
template <class T> tmp<gpuField<T>> operator-(const gpuField<T> &f)
{
    gpuField<T> *r = new gpuField<T>(f.size());
    for (label i = 0; i < f.size(); i++) r->data()[i] = -f.data()[i];
    return tmp<gpuField<T>>(r);
}

// ---- interface lists: the harness drives single-domain matrices, so no interface is ever set ----
template <template <class> class Field, class T> class FieldField
{
    label n_;
    std::vector<const Field<T> *> p_; // optional: entries pointed at caller fields (setPtr); unset entries read as empty

public:
    explicit FieldField(label n = 0) : n_(n), p_((size_t)n, nullptr) {}
    label size() const { return n_; }
    const Field<T> &operator[](label i) const
    {
        static Field<T> none;
        return (i < (label)p_.size() && p_[(size_t)i]) ? *p_[(size_t)i] : none;
    }
    void set(label, const tmp<Field<T>> &) {}
    void setPtr(label i, const Field<T> *f) { p_[(size_t)i] = f; }
};

#ifndef SHIM_REFERENCE_MATRIX_INTERFACES
class lduInterfaceFieldPtrsList
{
public:
    label size() const { return 0; }
    bool set(label) const { return false; }
};
#endif

// ---- ops.H (src/OpenFOAM/primitives/ops/ops.H:227-238 and the unary-operator functor macro) ----
template <class T> class unityOp
{
public:
    SHIM_HD T operator()(const T &x) const { return x; }
};
template <class T> class sumOp
{
public:
    SHIM_HD T operator()(const T &x, const T &y) const { return x + y; }
};
template <class R, class T> struct negateUnaryOperatorFunctor {
    SHIM_HD R operator()(const T &x) const { return -x; }
};

// ---- Textures.H (src/OpenFOAM/device/Textures.H): texture fetch == plain load ----
template <class T> class textures
{
    const T *data_;

public:
    SHIM_HD explicit textures(const T *d) : data_(d) {}
    SHIM_HD T operator[](const int &i) const { return data_[i]; }
};
template <class T> class textureBind
{
    const T *data_;

public:
    textureBind(const gpuList<T> &l) : data_(l.data()) {}
    textures<T> operator()() const { return textures<T>(data_); }
};

// ---- lduMatrixSolutionCache::favourSpeed (selects the reference's "fast" sorted-coefficient path) ----
struct lduMatrixSolutionCache {
    static int favourSpeed;
    static const gpuField<scalar> &first(label size); // scratch vectors (defined by the solver harness)
    static const gpuField<scalar> &second(label size);
};

// ---- lduAddressing: the arrays of LDU/lduAddressing/lduAddressing.H:200-255, supplied by the test ----
class lduAddressing
{
public:
    labelgpuList lower_, upper_, ownerStart_, losortStart_, losort_, ownerSort_;
    label nCells_;
    label size() const { return nCells_; }
    const labelgpuList &lowerAddr() const { return lower_; }
    const labelgpuList &upperAddr() const { return upper_; }
    const labelgpuList &ownerStartAddr() const { return ownerStart_; }
    const labelgpuList &losortStartAddr() const { return losortStart_; }
    const labelgpuList &losortAddr() const { return losort_; }
    const labelgpuList &ownerSortAddr() const { return ownerSort_; }
    // coupled-patch sort addressing of ONE patch (lduAddressing.C:38-167), supplied by the harness
    labelgpuList patchCells_, patchSort_, patchSortStart_;
    // ... or of every patch, by patch index, when the harness fills these
    std::vector<labelgpuList> patchCellsV_, patchSortV_, patchSortStartV_;
    const labelgpuList &patchSortCells(label p) const { return patchCellsV_.empty() ? patchCells_ : patchCellsV_[(size_t)p]; }
    const labelgpuList &patchSortAddr(label p) const { return patchSortV_.empty() ? patchSort_ : patchSortV_[(size_t)p]; }
    const labelgpuList &patchSortStartAddr(label p) const
    {
        return patchSortStartV_.empty() ? patchSortStart_ : patchSortStartV_[(size_t)p];
    }
};

// ---- lduMatrix: declarations of the members defined in the reference's lduMatrixATmul.C ----
class lduMatrix
{
public:
    class solver;         // defined in shim_solvers/solver_shim.h (only the solver harness needs them)
    class preconditioner;
    class smoother;
    static int debug;
    struct MeshStub {
        int comm() const { return 0; }
        template <class V, class Op> void reduce(V &, const Op &) const {} // one rank
    };
    MeshStub lduMesh_;
    const MeshStub &mesh() const { return lduMesh_; }
    bool diagonal() const { return diagPtr_ && !lowerPtr_ && !upperPtr_; } // lduMatrix.H:626-639
    bool symmetric() const { return diagPtr_ && (!lowerPtr_ && upperPtr_); }
    bool asymmetric() const { return diagPtr_ && lowerPtr_ && upperPtr_; }
    const lduAddressing *addr_;
    scalargpuField *lowerPtr_, *upperPtr_, *diagPtr_, *lowerSortPtr_, *upperSortPtr_;
    int level_;
    bool coarsest_;

    const lduAddressing &lduAddr() const { return *addr_; }
    const scalargpuField &lower() const { return lowerPtr_ ? *lowerPtr_ : *upperPtr_; }
    const scalargpuField &upper() const { return upperPtr_ ? *upperPtr_ : *lowerPtr_; }
    const scalargpuField &diag() const { return *diagPtr_; }
    const scalargpuField &lowerSort() const { return *lowerSortPtr_; }
    const scalargpuField &upperSort() const { return *upperSortPtr_; }
    bool coarsestLevel() const { return coarsest_; }
    int level() const { return level_; }
#ifdef SHIM_REFERENCE_MATRIX_OPERATIONS
    // defined by the reference (lduMatrixOperations.C:36-470)
    void sumDiag();
    void negSumDiag();
    void sumMagOffDiag(scalargpuField &) const;
    void operator=(const lduMatrix &);
    void negate();
    void operator+=(const lduMatrix &);
    void operator-=(const lduMatrix &);
    void operator*=(const scalargpuField &);
    void operator*=(scalar);
    // the allocate-on-demand accessors those operators rely on (lduMatrix.C:219-270): a missing triangle starts as a copy of
    // the other one, else as zeros; the harness owns (and leaks, per call) what they allocate
    scalargpuField &lower()
    {
        if (!lowerPtr_) lowerPtr_ = upperPtr_ ? new scalargpuField(*upperPtr_) : new scalargpuField(lduAddr().lowerAddr().size(), 0.0);
        lowerSortPtr_ = NULL;
        return *lowerPtr_;
    }
    scalargpuField &upper()
    {
        if (!upperPtr_) upperPtr_ = lowerPtr_ ? new scalargpuField(*lowerPtr_) : new scalargpuField(lduAddr().lowerAddr().size(), 0.0);
        upperSortPtr_ = NULL;
        return *upperPtr_;
    }
#endif
#ifdef SHIM_REFERENCE_MATRIX_INTERFACES
    // defined by the reference (lduMatrixUpdateMatrixInterfaces.C:30-276)
    void initMatrixInterfaces(const FieldField<gpuField, scalar> &, const lduInterfaceFieldPtrsList &,
                              const scalargpuField &, scalargpuField &, const direction, const bool negate = false) const;
    void updateMatrixInterfaces(const FieldField<gpuField, scalar> &, const lduInterfaceFieldPtrsList &,
                                const scalargpuField &, scalargpuField &, const direction, const bool negate = false) const;
    const lduSchedule &patchSchedule() const
    {
        static lduSchedule none;
        return none;
    }
#else
    void initMatrixInterfaces(const FieldField<gpuField, scalar> &, const lduInterfaceFieldPtrsList &,
                              const scalargpuField &, scalargpuField &, const direction) const
    {
    }
    void updateMatrixInterfaces(const FieldField<gpuField, scalar> &, const lduInterfaceFieldPtrsList &,
                                const scalargpuField &, scalargpuField &, const direction) const
    {
    }
    // the smoothers pass a sixth argument (negate): lduMatrixUpdateMatrixInterfaces.C:30-276
    void initMatrixInterfaces(const FieldField<gpuField, scalar> &, const lduInterfaceFieldPtrsList &,
                              const scalargpuField &, scalargpuField &, const direction, bool) const
    {
    }
    void updateMatrixInterfaces(const FieldField<gpuField, scalar> &, const lduInterfaceFieldPtrsList &,
                                const scalargpuField &, scalargpuField &, const direction, bool) const
    {
    }

#endif

    // defined by the reference (lduMatrixATmul.C:183-554)
    void Amul(scalargpuField &, const tmp<scalargpuField> &, const FieldField<gpuField, scalar> &,
              const lduInterfaceFieldPtrsList &, const direction) const;
    void Tmul(scalargpuField &, const tmp<scalargpuField> &, const FieldField<gpuField, scalar> &,
              const lduInterfaceFieldPtrsList &, const direction) const;
    void sumA(scalargpuField &, const FieldField<gpuField, scalar> &, const lduInterfaceFieldPtrsList &) const;
    void residual(scalargpuField &, const scalargpuField &, const scalargpuField &,
                  const FieldField<gpuField, scalar> &, const lduInterfaceFieldPtrsList &, const direction) const;
    tmp<scalargpuField> residual(const scalargpuField &, const scalargpuField &,
                                 const FieldField<gpuField, scalar> &, const lduInterfaceFieldPtrsList &,
                                 const direction) const;
    void H1(scalargpuField &) const;
    tmp<scalargpuField> H1() const;
    // defined by the reference (lduMatrixTemplates.C:50-160)
    template <class Type> void H(gpuField<Type> &, const gpuField<Type> &) const;
    template <class Type> tmp<gpuField<Type>> H(const gpuField<Type> &) const;
    template <class Type> tmp<gpuField<Type>> H(const tmp<gpuField<Type>> &) const;
    template <class Type> void faceH(gpuField<Type> &, const gpuField<Type> &) const;
    template <class Type> tmp<gpuField<Type>> faceH(const gpuField<Type> &) const;
    template <class Type> tmp<gpuField<Type>> faceH(const tmp<gpuField<Type>> &) const;
    scalargpuField &diag() // lduMatrix.C:237-245
    {
        if (!diagPtr_) diagPtr_ = new scalargpuField(lduAddr().size(), 0.0);
        return *diagPtr_;
    }
};

template <class T> struct pTraits;
template <> struct pTraits<scalar> {
    static constexpr scalar zero = 0.0, one = 1.0;
    enum { nComponents = 1 };
    static constexpr const char *componentNames[1] = {""};
};
struct FatalStream {
    std::string msg;
    template <class T> FatalStream &operator<<(const T &) { return *this; }
    FatalStream &operator<<(const char *c)
    {
        msg += c;
        return *this;
    }
};
static FatalStream FatalError;
#define FatalErrorIn(where) (::Foam::FatalError.msg.clear(), ::Foam::FatalError)
inline int exit(FatalStream &) { throw std::runtime_error("FatalError"); }
inline int abort(FatalStream &e) { throw std::runtime_error("FatalError: " + e.msg); } // error.H: abort(FatalError)
template <class R, class T> struct magUnaryFunctionFunctor {
    SHIM_HD R operator()(const T &x) const { return x < 0 ? -x : x; }
};
} // namespace Foam
#endif
