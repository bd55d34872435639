/*
 * containers.h -- host stand-ins for the OpenFOAM containers the reference's addressing sources use
 * (List, gpuList, PtrList, tmp, autoPtr, the error streams).  TEST INFRASTRUCTURE ONLY; shared by the
 * GAMGAgglomeration and lduAddressing shims.
 */
#ifndef SHIM_CONTAINERS_H
#define SHIM_CONTAINERS_H
#include <cstdlib>
#include <functional>
#include <memory>
#include <stdexcept>
#include <vector>

#include <thrust/copy.h>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/iterator/zip_iterator.h>
#include <thrust/transform.h>
#include <thrust/tuple.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/permutation_iterator.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#endif
#define __HOST____DEVICE__
// the "atomic" functor variants of the reference headers are compiled but never run here
template <class T> inline T atomicAdd(T *p, T v)
{
    T old = *p;
    *p += v;
    return old;
}

#define forAll(list, i) for (Foam::label i = 0; i < (list).size(); i++)
#define forAllReverse(list, i) for (Foam::label i = (list).size() - 1; i >= 0; i--)

namespace Foam
{
typedef double scalar;
typedef int label;

struct Istream; // no streams in the harnesses: the Istream constructors of the reference compile and throw

template <class T> class List
{
protected:
    std::vector<T> v_;

public:
    typedef T *iterator;
    typedef const T *const_iterator;
    List() {}
    explicit List(label n) : v_((size_t)n) {}
    List(label n, const T &x) : v_((size_t)n, x) {}
    List(const T *p, label n) : v_(p, p + n) {}
    explicit List(Istream &) { throw std::runtime_error("no streams in the harness"); }
    template <class DL> void transfer(DL &dl) // List::transfer(DynamicList&): take the contents over
    {
        v_.assign(dl.begin(), dl.end());
        dl.clear();
    }
    label size() const { return (label)v_.size(); }
    void setSize(label n) { v_.resize((size_t)n); }
    T &operator[](label i) { return v_[(size_t)i]; }
    const T &operator[](label i) const { return v_[(size_t)i]; }
    void operator=(const T &x)
    {
        for (auto &e : v_) e = x;
    }
    void set(label i, const T &x) { v_[(size_t)i] = x; }
    T *data() { return v_.data(); }
    const T *data() const { return v_.data(); }
    iterator begin() { return v_.data(); }
    iterator end() { return v_.data() + v_.size(); }
    const_iterator begin() const { return v_.data(); }
    const_iterator end() const { return v_.data() + v_.size(); }
};
template <class T> class gpuList : public List<T>
{
public:
    using List<T>::List;
    using List<T>::operator=;
    gpuList() {}
    gpuList(const List<T> &l) : List<T>(l) {}
    gpuList &operator=(const List<T> &l)
    {
        List<T>::v_.assign(l.begin(), l.end());
        return *this;
    }
};
typedef List<label> labelList, labelUList, labelField;
typedef List<unsigned char> boolList; // one byte per flag (std::vector<bool> has no addressable elements)
typedef List<labelList> labelListList;
typedef gpuList<label> labelgpuList, labelgpuField;
typedef gpuList<unsigned char> boolgpuList;
typedef List<labelgpuList> labelgpuListList;

template <class T> struct pTraits;
template <> struct pTraits<scalar> {
    static constexpr scalar zero = 0.0;
};

inline label min(const labelField &f)
{
    label m = f.size() ? f[0] : 0;
    forAll(f, i) if (f[i] < m) m = f[i];
    return m;
}

struct NullStream {
    template <class T> NullStream &operator<<(const T &) { return *this; }
};
static NullStream Pout, FatalError;
static const char endl = '\n';
#define FatalErrorIn(where) ::Foam::FatalError
inline int exit(NullStream &) { throw std::runtime_error("FatalError"); }
inline int abort(NullStream &) { throw std::runtime_error("FatalError"); }

struct Pstream {
    enum commsTypes { blocking, scheduled, nonBlocking };
    static bool parRun() { return false; }
    static void waitRequests() {}
};

template <class T> class tmp
{
    mutable T *p_;

public:
    tmp(T *p = nullptr) : p_(p) {}
    tmp(const tmp &o) : p_(o.p_) { o.p_ = nullptr; }
    ~tmp() { delete p_; }
    T &operator()() const { return *p_; }
    operator const T &() const { return *p_; }
};
template <class T> class autoPtr
{
    mutable T *p_;

public:
    autoPtr(T *p = nullptr) : p_(p) {}
    autoPtr(const autoPtr &o) : p_(o.p_) { o.p_ = nullptr; }
    ~autoPtr() { delete p_; }
    T *ptr() const
    {
        T *r = p_;
        p_ = nullptr;
        return r;
    }
};

template <class T> class PtrList
{
    std::vector<std::unique_ptr<T>> v_;

public:
    PtrList() {}
    explicit PtrList(label n) : v_((size_t)n) {}
    void setSize(label n) { v_.resize((size_t)n); }
    void clear() { v_.clear(); }
    label size() const { return (label)v_.size(); }
    bool set(label i) const { return v_[(size_t)i] != nullptr; }
    // Foam::PtrList::set returns the previous entry as an autoPtr (used to move a level, :764)
    autoPtr<T> set(label i, T *p)
    {
        T *old = v_[(size_t)i].release();
        v_[(size_t)i].reset(p);
        return autoPtr<T>(old);
    }
    autoPtr<T> set(label i, const autoPtr<T> &p) { return set(i, p.ptr()); }
    autoPtr<T> set(label i, const tmp<T> &t) { return set(i, new T(t())); }
    T &operator[](label i) const { return *v_[(size_t)i]; }
};

} // namespace Foam
#endif
