/*
 * solver_shim.h -- what the reference's solver sources (PCG.C, PBiCG.C, PBiCGStab.C,
 * AINVPreconditioner.C, diagonalPreconditioner.C) need around them to compile for the host:
 * words and a flat controls "dictionary", what SolverPerformance.H / .C name around them (those two files are the
 * reference's, included by path: convergence and singularity tests), the class declarations of lduMatrix::solver
 * (defined by the reference's lduMatrixSolver.C: New with its run-time selection tables, readControls,
 * normFactor) and ::preconditioner / ::smoother, a miniature of the run-time selection tables, the scratch-vector cache, serial global sums.  TEST INFRASTRUCTURE
 * ONLY.  The iteration loops, the order of their operations and their loop conditions are the
 * reference's (included by path); what is restated HERE, because the reference's versions live in
 * files tied to its I/O and run-time-selection machinery, is:
 *   gSumProd / gSumMag / gAverage as index-order serial sums  gpuFieldCommonFunctions.C:420-636
 */
#ifndef SOLVER_SHIM_H
#define SOLVER_SHIM_H
#include "foam_shim.h"

#include <cmath>
#include <map>
#include <stdexcept>
#include <memory>
#include <string>
#include <thrust/copy.h>

namespace Foam
{
class word : public std::string
{
public:
    word() {}
    word(const char *s) : std::string(s) {}
    word(const std::string &s) : std::string(s) {}
};
inline word operator+(const word &a, const word &b) { return word(static_cast<const std::string &>(a) + b); }
inline void operator>>(const word &in, word &out) { out = in; } // `dict.lookup("key") >> name`

struct dictionary;
typedef struct dictionary dictionaryFwd;
struct dictionary { // the keys the solver sources read (lduMatrixSolver.C:167-173, smoothSolver.C:80, JacobiSmoother.C:36)
    word solver, preconditioner, smoother;
    scalar tolerance = 1e-6, relTol = 0, omega = -1; // omega < 0: entry absent
    label maxIter = 1000, minIter = 0, nSweeps = 1;
    bool hasTolerance = true, hasRelTol = true, hasMaxIter = true, hasMinIter = true; // absent => the reference's defaults
    word lookup(const char *key) const
    {
        const std::string k(key);
        if (k == "solver") return solver;
        if (k == "preconditioner") return preconditioner;
        if (k == "smoother") return smoother;
        throw std::runtime_error("keyword " + k + " is undefined in dictionary");
    }
    // dictionary::lookupEntry(...): the word behind `preconditioner` / `smoother` (always a primitive entry here)
    struct entryStream {
        word w;
        void operator>>(word &out) const { out = w; }
    };
    struct entry {
        word w;
        bool isDict() const { return false; }
        const dictionary &dict() const { return dictionary::null; }
        entryStream stream() const { return entryStream{w}; }
    };
    static const dictionary null;
    entry lookupEntry(const char *key, bool, bool) const { return entry{lookup(key)}; }
    bool readIfPresent(const char *key, scalar &v) const
    {
        if (std::string(key) == "omega" && omega >= 0) {
            v = omega;
            return true;
        }
        return false;
    }
    template <class T> T lookupOrDefault(const char *key, const T &dflt) const
    {
        const std::string k(key);
        if (k == "nSweeps") return (T)nSweeps;
        if (k == "maxIter") return hasMaxIter ? (T)maxIter : dflt;
        if (k == "minIter") return hasMinIter ? (T)minIter : dflt;
        if (k == "tolerance") return hasTolerance ? (T)tolerance : dflt;
        if (k == "relTol") return hasRelTol ? (T)relTol : dflt;
        return dflt;
    }
};

typedef dictionary::entry entry;

#define defineTypeNameAndDebug(Type, DebugSwitch)         \
    const ::Foam::word Type::typeName(Type::typeName_()); \
    int Type::debug(DebugSwitch) /* className.H: the name comes from the class's TypeName("...") */

// ---- runTimeSelectionTables.H:49-120 in miniature: name -> constructor function, filled by the static
// add...ConstructorToTable objects the reference's .C files define ----
#define declareShimSelectionTable(baseType, argNames, argList, parList)                                          \
    typedef autoPtr<baseType>(*argNames##ConstructorPtr) argList;                                                \
    class argNames##ConstructorTable : public std::map<std::string, argNames##ConstructorPtr>                  \
    {                                                                                                            \
    public:                                                                                                      \
        struct iterator : std::map<std::string, argNames##ConstructorPtr>::iterator {                           \
            iterator(std::map<std::string, argNames##ConstructorPtr>::iterator i)                                \
                : std::map<std::string, argNames##ConstructorPtr>::iterator(i)                                   \
            {                                                                                                    \
            }                                                                                                    \
            argNames##ConstructorPtr operator()() const { return (*this)->second; }                             \
        };                                                                                                       \
        iterator find(const word &k) { return iterator(std::map<std::string, argNames##ConstructorPtr>::find(k)); } \
        iterator end() { return iterator(std::map<std::string, argNames##ConstructorPtr>::end()); }             \
        word sortedToc()                                                                                         \
        {                                                                                                        \
            std::string t;                                                                                       \
            for (auto i = std::map<std::string, argNames##ConstructorPtr>::begin();                              \
                 i != std::map<std::string, argNames##ConstructorPtr>::end(); ++i)                               \
                t += i->first + " ";                                                                             \
            return word(t);                                                                                      \
        }                                                                                                        \
    };                                                                                                           \
    static argNames##ConstructorTable *argNames##ConstructorTablePtr_;                                           \
    template <class T> class add##argNames##ConstructorToTable                                                   \
    {                                                                                                            \
    public:                                                                                                      \
        static autoPtr<baseType> New argList { return autoPtr<baseType>(new T parList); }                        \
        add##argNames##ConstructorToTable(const word &lookup = T::typeName)                                      \
        {                                                                                                        \
            if (!argNames##ConstructorTablePtr_) argNames##ConstructorTablePtr_ = new argNames##ConstructorTable; \
            (*argNames##ConstructorTablePtr_)[lookup] = New;                                                     \
        }                                                                                                        \
    }
#define SHIM_SOLVER_ARGS                                                                                         \
    (const word &f, const lduMatrix &m, const FieldField<gpuField, scalar> &b, const FieldField<gpuField, scalar> &i, \
     const lduInterfaceFieldPtrsList &l, const dictionary &d)
#define SHIM_SOLVER_PARS (f, m, b, i, l, d)
#define defineRunTimeSelectionTable(baseType, argNames) \
    baseType::argNames##ConstructorTable *baseType::argNames##ConstructorTablePtr_ = nullptr

struct NullStream {
    template <class T> NullStream &operator<<(const T &) { return *this; }
    NullStream &masterStream(int) { return *this; }
};
static NullStream Info;
static const char endl = '\n';
static const char nl = '\n';
struct FatalIOStream {
    std::string msg;
    template <class T> FatalIOStream &operator<<(const T &) { return *this; }
    FatalIOStream &operator<<(const word &w)
    {
        msg += w + " ";
        return *this;
    }
    FatalIOStream &operator<<(const char *c)
    {
        msg += c;
        return *this;
    }
};
static FatalIOStream FatalIOError;
#define FatalIOErrorIn(where, dict) (::Foam::FatalIOError.msg.clear(), ::Foam::FatalIOError)
inline int exit(FatalIOStream &e) { throw std::runtime_error(e.msg); }
struct Pstream {
    static int msgType() { return 0; }
};
template <class T, class Op> inline void reduce(T &, const Op &, int, int) {} // one rank

inline scalar mag(scalar x) { return std::fabs(x); }
template <class T> struct minusOp {
    T operator()(const T &x, const T &y) const { return x - y; }
};
template <class R, class A, class B> struct multiplyOperatorFunctor {
    R operator()(const A &a, const B &b) const { return a * b; }
};
template <class R, class S, class T> struct divideOperatorSFFunctor { // s / x
    const S s;
    divideOperatorSFFunctor(S _s) : s(_s) {}
    R operator()(const T &x) const { return s / x; }
};

template <class T> class autoPtr
{
    mutable T *p_;

public:
    autoPtr(T *p = nullptr) : p_(p) {}
    autoPtr(const autoPtr &o) : p_(o.p_) { o.p_ = nullptr; }
    autoPtr &operator=(const autoPtr &o) // transfers ownership, like Foam::autoPtr
    {
        if (this != &o) {
            delete p_;
            p_ = o.p_;
            o.p_ = nullptr;
        }
        return *this;
    }
    ~autoPtr() { delete p_; }
    T *operator->() const { return p_; }
    T &operator()() const { return *p_; }
    T *ptr() const
    {
        T *r = p_;
        p_ = nullptr;
        return r;
    }
};

// global sums on one rank, in index order (the oracle's order; the reference's thrust::reduce order is
// unspecified).  REF_OMP (the all-core timing build, oracle/Makefile `ref`): OpenMP reductions instead.
#ifdef REF_OMP
#define REF_REDUCE(var) _Pragma("omp parallel for schedule(static) reduction(+ : s)")
#else
#define REF_REDUCE(var)
#endif
inline scalar gSumMag(const scalargpuField &f, int)
{
    scalar s = 0;
    const scalar *p = f.data();
    const label n = f.size();
    REF_REDUCE(s)
    for (label i = 0; i < n; i++) s += std::fabs(p[i]);
    return s;
}
inline scalar gSumProd(const scalargpuField &a, const scalargpuField &b, int)
{
    scalar s = 0;
    const scalar *p = a.data(), *q = b.data();
    const label n = a.size();
    REF_REDUCE(s)
    for (label i = 0; i < n; i++) s += p[i] * q[i];
    return s;
}
inline scalar gAverage(const scalargpuField &f, int)
{
    scalar s = 0;
    const scalar *p = f.data();
    const label n = f.size();
    REF_REDUCE(s)
    for (label i = 0; i < n; i++) s += p[i];
    return s / f.size();
}

// ---- SolverPerformance<Type>: the REFERENCE'S OWN SolverPerformance.H / .C (checkConvergence, checkSingularity,
// singular, max, print) are included below by path; this block is what those two files name around them ----
template <class T, unsigned N> class FixedList
{
    T v_[N];

public:
    FixedList() {}
    FixedList(const T &x)
    {
        for (unsigned i = 0; i < N; i++) v_[i] = x;
    }
    T &operator[](label i) { return v_[i]; }
    const T &operator[](label i) const { return v_[i]; }
    bool operator!=(const FixedList &o) const
    {
        for (unsigned i = 0; i < N; i++)
            if (v_[i] != o.v_[i]) return true;
        return false;
    }
};
typedef NullStream Ostream;
struct Istream {
    void readBeginList(const char *) {}
    void readEndList(const char *) {}
    template <class T> Istream &operator>>(T &) { return *this; }
};
namespace token
{
enum punctuationToken { BEGIN_LIST = '(', END_LIST = ')', SPACE = ' ' };
}
inline scalar component(const scalar &s, const direction) { return s; }
inline scalar cmptMultiply(const scalar &a, const scalar &b) { return a * b; }
using std::max;
static const scalar VSMALL = 1e-300; // doubleScalar.H
#define ClassName(name)                             \
    static const char *typeName_() { return name; } \
    static const ::Foam::word typeName;             \
    static int debug
#define TypeName(name) /* typeInfo.H:57-59 */ \
    ClassName(name);                         \
    virtual const ::Foam::word &type() const { return typeName; }
#define defineNamedTemplateTypeNameAndDebug(Type, DebugSwitch) \
    template <> const ::Foam::word Type::typeName(Type::typeName_()); \
    template <> int Type::debug(DebugSwitch)
} // namespace Foam
#define NoRepository
#include "SolverPerformance.H" /* reference: class declaration; pulls SolverPerformance.C */
namespace Foam
{
typedef SolverPerformance<scalar> solverPerformance;          // solverPerformance.H:40-44
makeSolverPerformance(scalar); /* the reference's macro (SolverPerformance.H:261-276, used as in solverPerformance.C:31):
                                  great_ 1e20, small_ 1e-20, vsmall_ VSMALL */

// ---- lduMatrix::solver (lduMatrix.H:100-260), lduMatrix::preconditioner (:420-520) ----
class lduMatrix::solver // lduMatrix.H:100-260; its member functions are the reference's lduMatrixSolver.C
{
protected:
    word fieldName_;
    const lduMatrix &matrix_;
    const FieldField<gpuField, scalar> &interfaceBouCoeffs_;
    const FieldField<gpuField, scalar> &interfaceIntCoeffs_;
    lduInterfaceFieldPtrsList interfaces_;
    dictionary controlDict_;
    label maxIter_, minIter_;
    scalar tolerance_, relTol_;
    virtual void readControls();

public:
    declareShimSelectionTable(solver, symMatrix, SHIM_SOLVER_ARGS, SHIM_SOLVER_PARS);
    declareShimSelectionTable(solver, asymMatrix, SHIM_SOLVER_ARGS, SHIM_SOLVER_PARS);
    solver(const word &fieldName, const lduMatrix &matrix, const FieldField<gpuField, scalar> &interfaceBouCoeffs,
           const FieldField<gpuField, scalar> &interfaceIntCoeffs, const lduInterfaceFieldPtrsList &interfaces,
           const dictionary &solverControls);
    static autoPtr<solver> New(const word &fieldName, const lduMatrix &matrix,
                               const FieldField<gpuField, scalar> &interfaceBouCoeffs,
                               const FieldField<gpuField, scalar> &interfaceIntCoeffs,
                               const lduInterfaceFieldPtrsList &interfaces, const dictionary &solverControls);
    virtual ~solver() {}
    const lduMatrix &matrix() const { return matrix_; }
    virtual void read(const dictionary &);
    virtual solverPerformance solve(scalargpuField &psi, const scalargpuField &source, const direction cmpt = 0) const = 0;
    scalar normFactor(const scalargpuField &psi, const scalargpuField &source, const scalargpuField &Apsi,
                      scalargpuField &tmpField) const;
};

class lduMatrix::preconditioner // lduMatrix.H:420-520; New / getName are the reference's lduMatrixPreconditioner.C
{
protected:
    const solver &solver_;

public:
    declareShimSelectionTable(preconditioner, symMatrix, (const solver &sol, const dictionary &d), (sol, d));
    declareShimSelectionTable(preconditioner, asymMatrix, (const solver &sol, const dictionary &d), (sol, d));
    preconditioner(const solver &sol) : solver_(sol) {}
    virtual ~preconditioner() {}
    virtual void precondition(scalargpuField &wA, const scalargpuField &rA, const direction cmpt = 0) const = 0;
    virtual void preconditionT(scalargpuField &wT, const scalargpuField &rT, const direction cmpt = 0) const
    {
        precondition(wT, rT, cmpt); // lduMatrix.H:510-520 default
    }
    static word getName(const dictionary &solverControls);
    static autoPtr<preconditioner> New(const solver &sol, const dictionary &solverControls);
};

class lduMatrix::smoother // lduMatrix.H:262-414; New / getName are the reference's lduMatrixSmoother.C
{
protected:
    word fieldName_;
    const lduMatrix &matrix_;
    const FieldField<gpuField, scalar> &interfaceBouCoeffs_;
    const FieldField<gpuField, scalar> &interfaceIntCoeffs_;
    lduInterfaceFieldPtrsList interfaces_;

public:
    declareShimSelectionTable(smoother, symMatrix, SHIM_SOLVER_ARGS, SHIM_SOLVER_PARS);
    declareShimSelectionTable(smoother, asymMatrix, SHIM_SOLVER_ARGS, SHIM_SOLVER_PARS);
    smoother(const word &fieldName, const lduMatrix &matrix, const FieldField<gpuField, scalar> &bou,
             const FieldField<gpuField, scalar> &intc, const lduInterfaceFieldPtrsList &ifs);
    virtual ~smoother() {}
    virtual void smooth(scalargpuField &psi, const scalargpuField &source, const direction cmpt,
                        const label nSweeps) const = 0;
    static word getName(const dictionary &);
    static autoPtr<smoother> New(const word &fieldName, const lduMatrix &matrix, const FieldField<gpuField, scalar> &bou,
                                 const FieldField<gpuField, scalar> &intc, const lduInterfaceFieldPtrsList &ifs,
                                 const dictionary &solverControls);
};

// ---- scratch vectors: PCGCache.H / lduMatrixSolutionCache.H hand out process-lifetime buffers ----
struct ScratchPool {
    static const scalargpuField &get(const char *name, label size)
    {
        static std::map<std::string, std::unique_ptr<scalargpuField>> pool;
        auto &slot = pool[name];
        if (!slot || slot->size() < size) slot.reset(new scalargpuField(size));
        return *slot;
    }
};
struct PCGCache {
#define CACHE_FIELD(n) \
    static const scalargpuField &n(label, label size) { return ScratchPool::get(#n, size); }
    CACHE_FIELD(pA)
    CACHE_FIELD(wA)
    CACHE_FIELD(rA)
    CACHE_FIELD(pT)
    CACHE_FIELD(wT)
    CACHE_FIELD(rT)
    CACHE_FIELD(tA)
    CACHE_FIELD(result1)
#undef CACHE_FIELD
};
} // namespace Foam
#endif
