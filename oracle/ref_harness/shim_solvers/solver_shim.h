/*
 * solver_shim.h -- what the reference's solver sources (PCG.C, PBiCG.C, PBiCGStab.C,
 * AINVPreconditioner.C, diagonalPreconditioner.C) need around them to compile for the host:
 * words and a flat controls "dictionary", solverPerformance, the lduMatrix::solver /
 * ::preconditioner base classes, the scratch-vector cache, serial global sums.  TEST INFRASTRUCTURE
 * ONLY.  The iteration loops, the order of their operations and their loop conditions are the
 * reference's (included by path); what is restated HERE, because the reference's versions live in
 * files tied to its I/O and run-time-selection machinery, is:
 *   solverPerformance::checkConvergence / checkSingularity   SolverPerformance.C:32-43, 74-85
 *   lduMatrix::solver::normFactor                             lduMatrixSolver.C:205-236
 *   preconditioner selection incl. the DIC/DILU -> AINV alias lduMatrixPreconditioner.C:38-62
 *   gSumProd / gSumMag / gAverage as index-order serial sums  gpuFieldCommonFunctions.C:420-636
 */
#ifndef SOLVER_SHIM_H
#define SOLVER_SHIM_H
#include "foam_shim.h"

#include <cmath>
#include <map>
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

struct dictionary { // the keys lduMatrix::solver::readControls reads (lduMatrixSolver.C:167-173) + smoothers'
    word preconditioner, smoother;
    // GAMGSolver::readControls keys (GAMGSolver.C:209-249) are set on the solver object by the harness
    scalar tolerance = 1e-6, relTol = 0, omega = -1; // omega < 0: entry absent
    label maxIter = 1000, minIter = 0, nSweeps = 1;
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
        if (std::string(key) == "nSweeps") return (T)nSweeps;
        return dflt;
    }
};

#define defineTypeNameAndDebug(Type, DebugSwitch)         \
    const ::Foam::word Type::typeName(Type::typeName_()); \
    int Type::debug(DebugSwitch) /* className.H: the name comes from the class's TypeName("...") */

struct NullStream {
    template <class T> NullStream &operator<<(const T &) { return *this; }
    NullStream &masterStream(int) { return *this; }
};
static NullStream Info;
static const char endl = '\n';

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

class solverPerformance // SolverPerformance.H/.C
{
    word solverName_, fieldName_;
    scalar initialResidual_, finalResidual_;
    label noIterations_;
    bool converged_, singular_;

public:
    static constexpr scalar great_ = 1e20, small_ = 1e-20, vsmall_ = 1e-300; // SolverPerformance.H:269-275
    solverPerformance() : initialResidual_(0), finalResidual_(0), noIterations_(0), converged_(false), singular_(false) {}
    solverPerformance(const word &s, const word &f)
        : solverName_(s), fieldName_(f), initialResidual_(0), finalResidual_(0), noIterations_(0), converged_(false),
          singular_(false)
    {
    }
    const word &solverName() const { return solverName_; }
    scalar &initialResidual() { return initialResidual_; }
    scalar &finalResidual() { return finalResidual_; }
    label &nIterations() { return noIterations_; }
    bool converged() const { return converged_; }
    bool singular() const { return singular_; }
    bool checkConvergence(const scalar Tolerance, const scalar RelTolerance) // SolverPerformance.C:74-85
    {
        if (finalResidual_ < Tolerance || (RelTolerance > small_ && finalResidual_ < RelTolerance * initialResidual_))
            converged_ = true;
        else
            converged_ = false;
        return converged_;
    }
    template <class S> void print(S &) const {}
    bool checkSingularity(const scalar residual) // SolverPerformance.C:32-43
    {
        singular_ = residual < vsmall_;
        return singular_;
    }
};

// ---- lduMatrix::solver (lduMatrix.H:100-260), lduMatrix::preconditioner (:420-520) ----
class lduMatrix::solver
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

public:
    template <class T> struct addsymMatrixConstructorToTable {
    };
    template <class T> struct addasymMatrixConstructorToTable {
    };
    solver(const word &fieldName, const lduMatrix &matrix, const FieldField<gpuField, scalar> &bou,
           const FieldField<gpuField, scalar> &intc, const lduInterfaceFieldPtrsList &ifs, const dictionary &d)
        : fieldName_(fieldName), matrix_(matrix), interfaceBouCoeffs_(bou), interfaceIntCoeffs_(intc), interfaces_(ifs),
          controlDict_(d), maxIter_(d.maxIter), minIter_(d.minIter), tolerance_(d.tolerance), relTol_(d.relTol)
    {
    }
    virtual ~solver() {}
    const lduMatrix &matrix() const { return matrix_; }
    void readControls() // lduMatrixSolver.C:167-173
    {
        maxIter_ = controlDict_.maxIter;
        minIter_ = controlDict_.minIter;
        tolerance_ = controlDict_.tolerance;
        relTol_ = controlDict_.relTol;
    }
    virtual solverPerformance solve(scalargpuField &psi, const scalargpuField &source, const direction cmpt = 0) const = 0;
    // lduMatrixSolver.C:205-236 -- sumA is the reference's (lduMatrixATmul.C:345-395)
    scalar normFactor(const scalargpuField &psi, const scalargpuField &source, const scalargpuField &Apsi,
                      scalargpuField &tmpField) const
    {
        matrix_.sumA(tmpField, interfaceBouCoeffs_, interfaces_);
        const scalar average = gAverage(psi, 0);
        scalar s = 0;
        const scalar *t = tmpField.data(), *ap = Apsi.data(), *sr = source.data();
        const label n = psi.size();
        REF_REDUCE(s)
        for (label i = 0; i < n; i++) {
            const scalar tmpVal = average * t[i];
            s += std::fabs(ap[i] - tmpVal) + std::fabs(sr[i] - tmpVal);
        }
        return s + solverPerformance::small_;
    }
};

class lduMatrix::preconditioner
{
protected:
    const solver &solver_;

public:
    template <class T> struct addsymMatrixConstructorToTable {
    };
    template <class T> struct addasymMatrixConstructorToTable {
    };
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

class lduMatrix::smoother // lduMatrix.H:262-414
{
protected:
    word fieldName_;
    const lduMatrix &matrix_;
    const FieldField<gpuField, scalar> &interfaceBouCoeffs_;
    const FieldField<gpuField, scalar> &interfaceIntCoeffs_;
    lduInterfaceFieldPtrsList interfaces_;

public:
    template <class T> struct addsymMatrixConstructorToTable {
    };
    template <class T> struct addasymMatrixConstructorToTable {
    };
    smoother(const word &fieldName, const lduMatrix &matrix, const FieldField<gpuField, scalar> &bou,
             const FieldField<gpuField, scalar> &intc, const lduInterfaceFieldPtrsList &ifs)
        : fieldName_(fieldName), matrix_(matrix), interfaceBouCoeffs_(bou), interfaceIntCoeffs_(intc), interfaces_(ifs)
    {
    }
    virtual ~smoother() {}
    virtual void smooth(scalargpuField &psi, const scalargpuField &source, const direction cmpt,
                        const label nSweeps) const = 0;
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
