/*
 * ddt_shim.h -- what the reference's ddtScheme.C / EulerDdtScheme.C need on top of schemes_shim.h: the two class
 * declarations (ddtScheme.H:60-230, EulerDdtScheme.H:50-180), Time, and the GeometricField-level field algebra their
 * expressions are written in (one rounded operation per element and operator, internal field and every patch, as the
 * reference's GeometricFieldFunctions evaluate them).  Exercised: EulerDdtScheme<Type>::fvmDdt(vf) (:331-361),
 * fvcDdtPhiCorr(U, phi) (:523-551) and ddtScheme<Type>::fvcDdtPhiCoeff (ddtScheme.C:139-174).  TEST INFRASTRUCTURE ONLY.
 */
#ifndef DDT_SHIM_H
#define DDT_SHIM_H
#include "schemes_shim.h"

namespace Foam
{
template <class Type> struct flux { // fluxFieldType: scalar for a vector velocity (flux.H)
    typedef scalar type;
};
static const dimensionSet dimTime, dimLength, dimVelocity, dimArea;
inline dimensionedScalar operator/(scalar s, const dimensionedScalar &d) { return dimensionedScalar(s / d.value()); }
inline dimensionedScalar operator*(const dimensionedScalar &a, const dimensionedScalar &b) { return dimensionedScalar(a.value() * b.value()); }
// cell-field algebra named by the rho / alpha variants of the schemes, which are parsed but never run here
#define SHIM_INERT_VOL_OP(op)                                                                                          \
    template <class A, class B>                                                                                        \
    tmp<GeometricField<B, fvPatchField, volMesh>> operator op(const GeometricField<A, fvPatchField, volMesh> &,        \
                                                              const GeometricField<B, fvPatchField, volMesh> &)        \
    {                                                                                                                  \
        throw std::runtime_error("not used");                                                                          \
    }                                                                                                                  \
    template <class A, class B>                                                                                        \
    tmp<GeometricField<B, fvPatchField, volMesh>> operator op(const tmp<GeometricField<A, fvPatchField, volMesh>> &,   \
                                                              const GeometricField<B, fvPatchField, volMesh> &)        \
    {                                                                                                                  \
        throw std::runtime_error("not used");                                                                          \
    }                                                                                                                  \
    template <class A, class B>                                                                                        \
    typename GeometricField<B, fvPatchField, volMesh>::GeometricBoundaryField operator op(                             \
        const typename GeometricField<A, fvPatchField, volMesh>::GeometricBoundaryField &,                             \
        const typename GeometricField<B, fvPatchField, volMesh>::GeometricBoundaryField &)                             \
    {                                                                                                                  \
        throw std::runtime_error("not used");                                                                          \
    }
SHIM_INERT_VOL_OP(*)
SHIM_INERT_VOL_OP(-)
SHIM_INERT_VOL_OP(+)
#undef SHIM_INERT_VOL_OP

// ---- element-wise GeometricField algebra: internal field + patch by patch ----
template <class R, class A, class B, class Op>
tmp<GeometricField<R, fvsPatchField, surfaceMesh>> gfBinary(const GeometricField<A, fvsPatchField, surfaceMesh> &a,
                                                            const GeometricField<B, fvsPatchField, surfaceMesh> &b, Op op)
{
    GeometricField<R, fvsPatchField, surfaceMesh> *r = new GeometricField<R, fvsPatchField, surfaceMesh>;
    r->mesh_ = a.mesh_;
    r->internal_.setSize(a.internal_.size());
    for (label i = 0; i < a.internal_.size(); i++) r->internal_.data()[i] = op(a.internal_.data()[i], b.internal_.data()[i]);
    r->boundary_.p_.resize(a.boundary_.p_.size());
    for (size_t p = 0; p < a.boundary_.p_.size(); p++) {
        r->boundary_.p_[p].setSize(a.boundary_.p_[p].size());
        for (label i = 0; i < a.boundary_.p_[p].size(); i++)
            r->boundary_.p_[p].data()[i] = op(a.boundary_.p_[p].data()[i], b.boundary_.p_[p].data()[i]);
    }
    return tmp<GeometricField<R, fvsPatchField, surfaceMesh>>(r);
}
template <class R, class A, class Op>
tmp<GeometricField<R, fvsPatchField, surfaceMesh>> gfUnary(const GeometricField<A, fvsPatchField, surfaceMesh> &a, Op op)
{
    return gfBinary<R>(a, a, [op](const A &x, const A &) { return op(x); });
}
inline scalar dot3(const vector &a, const vector &b) { return (a.v_[0] * b.v_[0] + a.v_[1] * b.v_[1]) + a.v_[2] * b.v_[2]; } // VectorI.H
inline tmp<surfaceScalarField> operator&(const surfaceVectorField &a, const tmp<surfaceVectorField> &b)
{
    return gfBinary<scalar>(a, b(), [](const vector &x, const vector &y) { return dot3(x, y); });
}
inline tmp<surfaceScalarField> operator-(const surfaceScalarField &a, const tmp<surfaceScalarField> &b)
{
    return gfBinary<scalar>(a, b(), [](scalar x, scalar y) { return x - y; });
}
inline tmp<surfaceScalarField> mag(const surfaceScalarField &a) { return gfUnary<scalar>(a, [](scalar x) { return std::fabs(x); }); }
inline tmp<surfaceScalarField> operator+(const tmp<surfaceScalarField> &a, const dimensionedScalar &d)
{
    const scalar s = d.value();
    return gfUnary<scalar>(a(), [s](scalar x) { return x + s; });
}
inline tmp<surfaceScalarField> operator/(const tmp<surfaceScalarField> &a, const tmp<surfaceScalarField> &b)
{
    return gfBinary<scalar>(a(), b(), [](scalar x, scalar y) { return x / y; });
}
inline tmp<surfaceScalarField> min(const tmp<surfaceScalarField> &a, scalar s)
{
    return gfUnary<scalar>(a(), [s](scalar x) { return x < s ? x : s; });
}
inline tmp<surfaceScalarField> operator-(scalar s, const tmp<surfaceScalarField> &a)
{
    return gfUnary<scalar>(a(), [s](scalar x) { return s - x; });
}
inline tmp<surfaceScalarField> operator*(const tmp<surfaceScalarField> &a, const dimensionedScalar &d)
{
    const scalar s = d.value();
    return gfUnary<scalar>(a(), [s](scalar x) { return x * s; });
}
inline tmp<surfaceScalarField> operator*(const tmp<surfaceScalarField> &a, const surfaceScalarField &b)
{
    return gfBinary<scalar>(a(), b, [](scalar x, scalar y) { return x * y; });
}
inline scalar gAverage(const scalargpuField &) { return 0; }
inline scalar gMax(const scalargpuField &) { return 0; }
inline scalar gMin(const scalargpuField &) { return 0; }
inline tmp<gpuField<vector>> operator*(scalar s, const gpuField<vector> &b)
{
    gpuField<vector> *r = new gpuField<vector>(b.size());
    for (label i = 0; i < b.size(); i++) r->data()[i] = s * b.data()[i];
    return tmp<gpuField<vector>>(r);
}
inline tmp<gpuField<vector>> operator*(const tmp<gpuField<vector>> &a, const scalargpuField &b)
{
    gpuField<vector> *r = new gpuField<vector>(b.size());
    for (label i = 0; i < b.size(); i++) r->data()[i] = a().data()[i] * b.data()[i];
    return tmp<gpuField<vector>>(r);
}

namespace fvc
{
// fvc::interpolate with the linear scheme (surfaceInterpolate.C:300-316 -> scheme(mesh, name)().interpolate(vf)): the
// reference's surfaceInterpolationScheme<Type>::interpolate(vf) with the mesh's linear weights
template <class Type>
tmp<GeometricField<Type, fvsPatchField, surfaceMesh>> interpolate(const GeometricField<Type, fvPatchField, volMesh> &vf)
{
    return surfaceInterpolationScheme<Type>(vf.mesh()).interpolate(vf);
}
} // namespace fvc

namespace fv
{
static int debug = 0;
template <class Type> class ddtScheme : public refCount // ddtScheme.H:60-230
{
protected:
    const fvMesh &mesh_;

public:
    typedef GeometricField<typename flux<Type>::type, fvsPatchField, surfaceMesh> fluxFieldType;
    struct IstreamConstructorTable {
        struct iterator {
            bool operator==(const iterator &) const { return true; }
            tmp<ddtScheme<Type>> (*operator()() const)(const fvMesh &, Istream &) { return nullptr; }
        };
        iterator find(const word &) { return iterator(); }
        iterator end() { return iterator(); }
        word sortedToc() const { return word(); }
    };
    static IstreamConstructorTable *IstreamConstructorTablePtr_;
    ddtScheme(const fvMesh &mesh) : mesh_(mesh) {}
    static tmp<ddtScheme<Type>> New(const fvMesh &mesh, Istream &schemeData);
    virtual ~ddtScheme();
    const fvMesh &mesh() const { return mesh_; }
    /* not virtual here: only the members the harness calls are instantiated */
    tmp<GeometricField<Type, fvPatchField, volMesh>> fvcDdt(const volScalarField &alpha, const volScalarField &rho,
                                                                    const GeometricField<Type, fvPatchField, volMesh> &vf);
    tmp<fvMatrix<Type>> fvmDdt(const volScalarField &alpha, const volScalarField &rho,
                                       const GeometricField<Type, fvPatchField, volMesh> &vf);
    tmp<surfaceScalarField> fvcDdtPhiCoeff(const GeometricField<Type, fvPatchField, volMesh> &U, const fluxFieldType &phi,
                                           const fluxFieldType &phiCorr);
    tmp<surfaceScalarField> fvcDdtPhiCoeff(const GeometricField<Type, fvPatchField, volMesh> &U, const fluxFieldType &phi);
};
template <class Type> class EulerDdtScheme : public ddtScheme<Type> // EulerDdtScheme.H:50-180
{
public:
    EulerDdtScheme(const fvMesh &mesh) : ddtScheme<Type>(mesh) {}
    const fvMesh &mesh() const { return fv::ddtScheme<Type>::mesh(); }
    typedef typename ddtScheme<Type>::fluxFieldType fluxFieldType;
    tmp<GeometricField<Type, fvPatchField, volMesh>> fvcDdt(const dimensioned<Type> &);
    tmp<GeometricField<Type, fvPatchField, volMesh>> fvcDdt(const GeometricField<Type, fvPatchField, volMesh> &);
    tmp<GeometricField<Type, fvPatchField, volMesh>> fvcDdt(const dimensionedScalar &, const GeometricField<Type, fvPatchField, volMesh> &);
    tmp<GeometricField<Type, fvPatchField, volMesh>> fvcDdt(const volScalarField &, const GeometricField<Type, fvPatchField, volMesh> &);
    tmp<GeometricField<Type, fvPatchField, volMesh>> fvcDdt(const volScalarField &alpha, const volScalarField &rho,
                                                            const GeometricField<Type, fvPatchField, volMesh> &psi);
    tmp<fvMatrix<Type>> fvmDdt(const GeometricField<Type, fvPatchField, volMesh> &);
    tmp<fvMatrix<Type>> fvmDdt(const dimensionedScalar &, const GeometricField<Type, fvPatchField, volMesh> &);
    tmp<fvMatrix<Type>> fvmDdt(const volScalarField &, const GeometricField<Type, fvPatchField, volMesh> &);
    tmp<fvMatrix<Type>> fvmDdt(const volScalarField &alpha, const volScalarField &rho,
                               const GeometricField<Type, fvPatchField, volMesh> &psi);
    tmp<fluxFieldType> fvcDdtUfCorr(const GeometricField<Type, fvPatchField, volMesh> &U,
                                    const GeometricField<Type, fvsPatchField, surfaceMesh> &Uf);
    tmp<fluxFieldType> fvcDdtPhiCorr(const GeometricField<Type, fvPatchField, volMesh> &U, const fluxFieldType &phi);
    tmp<fluxFieldType> fvcDdtUfCorr(const volScalarField &rho, const GeometricField<Type, fvPatchField, volMesh> &U,
                                    const GeometricField<Type, fvsPatchField, surfaceMesh> &Uf);
    tmp<fluxFieldType> fvcDdtPhiCorr(const volScalarField &rho, const GeometricField<Type, fvPatchField, volMesh> &U,
                                     const fluxFieldType &phi);
    tmp<surfaceScalarField> meshPhi(const GeometricField<Type, fvPatchField, volMesh> &);
};
} // namespace fv
} // namespace Foam
#endif
