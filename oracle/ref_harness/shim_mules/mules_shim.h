/*
 * mules_shim.h -- the finite-volume types the reference's MULESTemplates.C is written against, reduced to what
 * MULES::limiter / limit / explicitSolve touch: an fvMesh with the LDU addressing, the per-patch sort addressing, the
 * cell volumes and a time step; cell and face fields with per-patch values; the gpuField algebra of the bound
 * expressions (one loop, one rounding per written operator, as gpuFieldFunctions evaluates them); a sliced face
 * field over one flat list.  The reference's own one / zero / oneField / zeroField / geometricOneField headers are
 * used as they are (symlinked by the Makefile).  Restated here instead of included: upwind<scalar>::flux
 * (upwind.H:86-103 weights = pos(faceFlux); surfaceInterpolationScheme.C:263-321 interpolate; :176-184 flux =
 * faceFlux*interpolate) -- pinned on its own by harness_fvm / harness_limiters.  TEST INFRASTRUCTURE ONLY.
 */
#ifndef MULES_SHIM_H
#define MULES_SHIM_H
#include "../shim/foam_shim.h"

#include <string>
#include <thrust/for_each.h>
#include <thrust/iterator/constant_iterator.h>

namespace Foam
{
static const scalar SMALL = 1e-15, VSMALL = 1e-300; /* doubleScalar.H */
inline scalar max(const scalar a, const scalar b) { return (a > b) ? a : b; }
inline scalar min(const scalar a, const scalar b) { return (a < b) ? a : b; }
struct Ostream {
    template <class T> Ostream &operator<<(const T &) { return *this; }
};
static Ostream Info;
static const char endl = '\n';
class word : public std::string
{
public:
    word() {}
    word(const char *s) : std::string(s) {}
};
struct dimensionSet {
};
static const dimensionSet dimless;
struct IOobject {
    enum readOption { NO_READ };
    enum writeOption { NO_WRITE };
    template <class... A> IOobject(const A &...) {}
};
template <class T> struct minOp {
};
typedef gpuField<scalar> scalarField; /* only named by limitSum's patch branch, never instantiated */
template <class T> class UPtrList
{
    std::vector<T *> p_;

public:
    explicit UPtrList(label n) : p_((size_t)n) {}
    label size() const { return (label)p_.size(); }
    template <class U> void set(label i, U *p) { p_[(size_t)i] = p; }
    T &operator[](label i) { return *p_[(size_t)i]; }
};

/* ---- scalargpuField algebra (gpuFieldFunctions: every operator is one pass with one rounding) ---- */
typedef tmp<scalargpuField> tsf;
#define MULES_BINOP(op)                                                                                                          \
    inline tsf operator op(const scalargpuField &a, const scalargpuField &b)                                                    \
    {                                                                                                                            \
        scalargpuField *r = new scalargpuField(a.size());                                                                        \
        for (label i = 0; i < a.size(); i++) r->data()[i] = a.data()[i] op b.data()[i];                                          \
        return tsf(r);                                                                                                           \
    }                                                                                                                            \
    inline tsf operator op(const tsf &a, const scalargpuField &b) { return a() op b; }                                           \
    inline tsf operator op(const scalargpuField &a, const tsf &b) { return a op b(); }                                           \
    inline tsf operator op(const tsf &a, const tsf &b) { return a() op b(); }                                                    \
    inline tsf operator op(const scalargpuField &a, const scalar &s)                                                             \
    {                                                                                                                            \
        scalargpuField *r = new scalargpuField(a.size());                                                                        \
        for (label i = 0; i < a.size(); i++) r->data()[i] = a.data()[i] op s;                                                    \
        return tsf(r);                                                                                                           \
    }                                                                                                                            \
    inline tsf operator op(const scalar &s, const scalargpuField &a)                                                             \
    {                                                                                                                            \
        scalargpuField *r = new scalargpuField(a.size());                                                                        \
        for (label i = 0; i < a.size(); i++) r->data()[i] = s op a.data()[i];                                                    \
        return tsf(r);                                                                                                           \
    }                                                                                                                            \
    inline tsf operator op(const tsf &a, const scalar &s) { return a() op s; }                                                   \
    inline tsf operator op(const scalar &s, const tsf &a) { return s op a(); }
MULES_BINOP(+)
MULES_BINOP(-)
MULES_BINOP(*)
MULES_BINOP(/)
#undef MULES_BINOP
inline tsf operator-(const tsf &a) { return -a(); }
inline tsf min(const scalargpuField &a, const scalar &s)
{
    scalargpuField *r = new scalargpuField(a.size());
    for (label i = 0; i < a.size(); i++) r->data()[i] = min(a.data()[i], s);
    return tsf(r);
}
inline tsf max(const scalargpuField &a, const scalar &s)
{
    scalargpuField *r = new scalargpuField(a.size());
    for (label i = 0; i < a.size(); i++) r->data()[i] = max(a.data()[i], s);
    return tsf(r);
}

/* ---- mesh ---- */
struct Time {
    scalar deltaT_ = 1;
    scalar deltaTValue() const { return deltaT_; }
    word timeName() const { return word("0"); }
};
struct objectRegistry {
    template <class T> const T &lookupObject(const word &) const { throw std::runtime_error("lookupObject"); }
};
struct fvPatch {
    labelgpuList faceCells_;
    virtual ~fvPatch() {}
    const labelgpuList &faceCells() const { return faceCells_; }
};
struct wedgeFvPatch : fvPatch {
};
template <class T, class U> inline bool isA(const U &u) { return dynamic_cast<const T *>(&u) != nullptr; }
struct DimensionedInternalField { /* volScalarField::DimensionedInternalField: V, Sp, Su */
    scalargpuField f_;
    const scalargpuField &getField() const { return f_; }
    operator const scalargpuField &() const { return f_; } /* CMULESTemplates.C:414: const scalargpuField& V = tVsc(); */
};
struct dictionary { /* mesh.solverDict(psi.name()): the two MULES controls CMULESTemplates.C reads */
    scalar extremaCoeff_ = 0;
    label nLimiterIter_ = 3;
    template <class T> T lookupOrDefault(const word &, const T &) const { return (T)extremaCoeff_; }
    label lookup(const word &) const { return nLimiterIter_; }
};
inline label readLabel(label l) { return l; }
class fvMesh : public objectRegistry
{
public:
    lduAddressing addr_;
    Time time_;
    DimensionedInternalField V_;
    label nInternalFaces_ = 0, nFaces_ = 0;
    std::vector<fvPatch> patches_;
    std::vector<label> patchStart_; /* into the boundary faces, one past the end appended */
    const lduAddressing &lduAddr() const { return addr_; }
    const labelgpuList &owner() const { return addr_.lowerAddr(); }
    const labelgpuList &neighbour() const { return addr_.upperAddr(); }
    bool moving() const { return false; }
    tmp<DimensionedInternalField> Vsc() const { return tmp<DimensionedInternalField>(V_); }
    tmp<DimensionedInternalField> Vsc0() const { return tmp<DimensionedInternalField>(V_); }
    const Time &time() const { return time_; }
    dictionary solverDict_;
    const dictionary &solverDict(const word &) const { return solverDict_; }
    label nFaces() const { return nFaces_; }
    label nInternalFaces() const { return nInternalFaces_; }
    const std::vector<fvPatch> &boundary() const { return patches_; }
};

/* ---- fields ---- */
class fvPatchScalarField : public scalargpuField
{
public:
    using scalargpuField::scalargpuField;
    using scalargpuField::operator=;
    bool coupled_ = false; /* processor / cyclic patch: the harness hands the neighbour values over as the patch values */
    bool coupled() const { return coupled_; }
    tsf patchNeighbourField() const { return tsf(new scalargpuField(static_cast<const scalargpuField &>(*this))); }
};
class fvsPatchScalarField : public scalargpuField
{
public:
    using scalargpuField::scalargpuField;
    using scalargpuField::operator=;
    fvsPatchScalarField() {}
    fvsPatchScalarField(const fvsPatchScalarField &o) : scalargpuField(static_cast<const scalargpuField &>(o)) {}
    fvsPatchScalarField &operator=(const fvsPatchScalarField &o)
    {
        scalargpuField::operator=(static_cast<const scalargpuField &>(o));
        return *this;
    }
    bool coupled() const { return false; }
};

class volScalarField : public scalargpuField
{
public:
    typedef std::vector<fvPatchScalarField> GeometricBoundaryField;
    typedef Foam::DimensionedInternalField DimensionedInternalField;
    const fvMesh *mesh_ = nullptr;
    const volScalarField *old_ = nullptr;
    GeometricBoundaryField boundary_;
    const fvMesh &mesh() const { return *mesh_; }
    const scalargpuField &getField() const { return *this; }
    scalargpuField &getField() { return *this; }
    const scalargpuField &internalField() const { return *this; }
    scalargpuField &internalField() { return *this; }
    const volScalarField &oldTime() const { return *old_; }
    const GeometricBoundaryField &boundaryField() const { return boundary_; }
    word name() const { return word("psi"); }
    void correctBoundaryConditions() {}
};

class surfaceScalarField : public scalargpuField
{
public:
    typedef std::vector<fvsPatchScalarField> GeometricBoundaryField;
    const fvMesh *mesh_ = nullptr;
    GeometricBoundaryField boundary_;
    surfaceScalarField() {}
    surfaceScalarField(const surfaceScalarField &o)
        : scalargpuField(static_cast<const scalargpuField &>(o)), mesh_(o.mesh_), boundary_(o.boundary_)
    {
    }
    surfaceScalarField(const tmp<surfaceScalarField> &t) : surfaceScalarField(t()) {}
    const fvMesh &mesh() const { return *mesh_; }
    const GeometricBoundaryField &boundaryField() const { return boundary_; }
    GeometricBoundaryField &boundaryField() { return boundary_; }
    /* GeometricField operators: the internal field, then every patch field */
    void operator-=(const surfaceScalarField &o)
    {
        scalargpuField::operator-=(o);
        for (size_t p = 0; p < boundary_.size(); p++) boundary_[p] -= o.boundary_[p];
    }
    void operator*=(const surfaceScalarField &o)
    {
        scalargpuField::operator*=(o);
        for (size_t p = 0; p < boundary_.size(); p++) boundary_[p] *= o.boundary_[p];
    }
    void operator=(const tmp<surfaceScalarField> &t)
    {
        scalargpuField::operator=(static_cast<const scalargpuField &>(t()));
        for (size_t p = 0; p < boundary_.size(); p++) boundary_[p] = t().boundary_[p];
    }
};
#define MULES_SURFOP(op)                                                                                                         \
    inline tmp<surfaceScalarField> operator op(const surfaceScalarField &a, const surfaceScalarField &b)                        \
    {                                                                                                                            \
        surfaceScalarField *r = new surfaceScalarField(a);                                                                       \
        for (label i = 0; i < a.size(); i++) r->data()[i] = a.data()[i] op b.data()[i];                                          \
        for (size_t p = 0; p < a.boundary_.size(); p++)                                                                          \
            for (label i = 0; i < a.boundary_[p].size(); i++)                                                                    \
                r->boundary_[p].data()[i] = a.boundary_[p].data()[i] op b.boundary_[p].data()[i];                                \
        return tmp<surfaceScalarField>(r);                                                                                       \
    }                                                                                                                            \
    inline tmp<surfaceScalarField> operator op(const surfaceScalarField &a, const tmp<surfaceScalarField> &b) { return a op b(); }
MULES_SURFOP(+)
MULES_SURFOP(*)
#undef MULES_SURFOP

/* slicedSurfaceScalarField(io, mesh, dims, completeField, preserveCouples): views into one flat face list */
class slicedSurfaceScalarField : public surfaceScalarField
{
public:
    slicedSurfaceScalarField(const IOobject &, const fvMesh &mesh, const dimensionSet &, const scalargpuField &all, bool)
    {
        mesh_ = &mesh;
        view(all.data(), mesh.nInternalFaces());
        boundary_.resize(mesh.patches_.size());
        for (size_t p = 0; p < boundary_.size(); p++)
            boundary_[p].view(all.data() + mesh.nInternalFaces() + mesh.patchStart_[p], mesh.patchStart_[p + 1] - mesh.patchStart_[p]);
    }
};

namespace syncTools
{
template <class M, class L, class Op> inline void syncFaceList(const M &, L &, const Op &) {} /* single domain */
}

template <class Type> class upwind
{
    const fvMesh &mesh_;
    const surfaceScalarField &faceFlux_;

public:
    upwind(const fvMesh &mesh, const surfaceScalarField &faceFlux) : mesh_(mesh), faceFlux_(faceFlux) {}
    tmp<surfaceScalarField> flux(const volScalarField &vf) const
    {
        surfaceScalarField *r = new surfaceScalarField(faceFlux_);
        const label *l = mesh_.owner().data(), *u = mesh_.neighbour().data();
        for (label f = 0; f < faceFlux_.size(); f++) {
            const scalar w = faceFlux_.data()[f] >= 0 ? 1.0 : 0.0; /* pos() */
            const scalar sf = w * (vf.data()[l[f]] - vf.data()[u[f]]) + vf.data()[u[f]];
            r->data()[f] = faceFlux_.data()[f] * sf;
        }
        for (size_t p = 0; p < r->boundary_.size(); p++)
            for (label i = 0; i < r->boundary_[p].size(); i++)
                r->boundary_[p].data()[i] = faceFlux_.boundary_[p].data()[i] * vf.boundary_[p].data()[i];
        return tmp<surfaceScalarField>(r);
    }
};

namespace fvc
{
/* fvc::surfaceIntegrate(ivf, ssf), fvcSurfaceIntegrate.C:136-205 (pinned by harness_fv.cpp; restated with the same order:
 * owner faces added, neighbour faces subtracted, boundary faces patch by patch, then / V) */
inline void surfaceIntegrate(scalargpuField &ivf, const surfaceScalarField &ssf)
{
    const fvMesh &mesh = ssf.mesh();
    const lduAddressing &a = mesh.lduAddr();
    for (label c = 0; c < ivf.size(); c++) {
        scalar s = ivf.data()[c];
        for (label f = a.ownerStart_.data()[c]; f < a.ownerStart_.data()[c + 1]; f++) s += ssf.data()[f];
        for (label k = a.losortStart_.data()[c]; k < a.losortStart_.data()[c + 1]; k++) s -= ssf.data()[a.losort_.data()[k]];
        ivf.data()[c] = s;
    }
    for (size_t p = 0; p < ssf.boundary_.size(); p++) {
        const labelgpuList &pc = a.patchSortCells((label)p), &ps = a.patchSortAddr((label)p), &pss = a.patchSortStartAddr((label)p);
        for (label i = 0; i < pc.size(); i++) {
            scalar s = ivf.data()[pc.data()[i]];
            for (label k = pss.data()[i]; k < pss.data()[i + 1]; k++) s += ssf.boundary_[p].data()[ps.data()[k]];
            ivf.data()[pc.data()[i]] = s;
        }
    }
    for (label c = 0; c < ivf.size(); c++) ivf.data()[c] /= mesh.V_.f_.data()[c];
}
} // namespace fvc
} // namespace Foam
#endif
