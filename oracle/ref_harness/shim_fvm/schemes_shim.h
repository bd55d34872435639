/*
 * schemes_shim.h -- class declarations and helpers the reference's gaussConvectionScheme.C / gaussLaplacianScheme.C are
 * written against (convectionScheme.H, laplacianScheme.H, surfaceInterpolationScheme.H reduced to what fvmDiv and
 * fvmLaplacianUncorrected touch), on top of fvm_shim.h.  The boundary-condition coefficient functions of the patch fields
 * (fixedValue, zeroGradient, coupled) are the three-line formulas of fixedValueFvPatchField.C:113-146,
 * zeroGradientFvPatchField.C:111-150 and coupledFvPatchField.C:162-209, restated in fvm_shim.h's fvPatchField; what the
 * reference's scheme sources contribute -- and what is being pinned -- is how the schemes combine them with the face flux,
 * gamma*magSf and the weights, and the coefficient fills themselves.  TEST INFRASTRUCTURE ONLY.
 */
#ifndef SCHEMES_SHIM_H
#define SCHEMES_SHIM_H
#include "fvm_shim.h"

namespace Foam
{
typedef GeometricField<vector, fvsPatchField, surfaceMesh> surfaceVectorField;
// surface-field algebra named by the non-orthogonal / tensor-gamma members, which are parsed but never run here
inline surfaceVectorField operator/(const surfaceVectorField &, const surfaceScalarField &) { throw std::runtime_error("not used"); }
inline surfaceScalarField operator&(const surfaceVectorField &, const surfaceVectorField &) { throw std::runtime_error("not used"); }
inline surfaceVectorField operator*(const surfaceScalarField &, const surfaceVectorField &) { throw std::runtime_error("not used"); }
inline surfaceVectorField operator-(const surfaceVectorField &, const surfaceVectorField &) { throw std::runtime_error("not used"); }
template <class T>
tmp<GeometricField<T, fvsPatchField, surfaceMesh>> operator*(const surfaceScalarField &, const tmp<GeometricField<T, fvsPatchField, surfaceMesh>> &)
{
    throw std::runtime_error("not used"); // faceFlux*correction(vf): only on the corrected() branch
}
// patch-field algebra of the coupled branch of surfaceInterpolationScheme.C:357-370 (one rounded operation per element)
inline tmp<scalargpuField> operator-(scalar s, const scalargpuField &b)
{
    scalargpuField *r = new scalargpuField(b.size());
    for (label i = 0; i < b.size(); i++) r->data()[i] = s - b.data()[i];
    return tmp<scalargpuField>(r);
}
inline tmp<gpuField<vector>> operator*(const scalargpuField &a, const tmp<gpuField<vector>> &b)
{
    gpuField<vector> *r = new gpuField<vector>(a.size());
    for (label i = 0; i < a.size(); i++) r->data()[i] = a.data()[i] * b().data()[i];
    return tmp<gpuField<vector>>(r);
}
inline tmp<gpuField<vector>> operator*(const tmp<scalargpuField> &a, const tmp<gpuField<vector>> &b) { return a() * b; }
inline tmp<gpuField<vector>> operator+(const tmp<gpuField<vector>> &a, const tmp<gpuField<vector>> &b)
{
    gpuField<vector> *r = new gpuField<vector>(a().size());
    for (label i = 0; i < a().size(); i++) r->data()[i] = a().data()[i] + b().data()[i];
    return tmp<gpuField<vector>>(r);
}
// Declaration for the reference's surfaceInterpolationScheme.C (FV/interpolation/surfaceInterpolation/
// surfaceInterpolationScheme/surfaceInterpolationScheme.H:56-260, the members its .C file defines); not virtual here, so only
// the members the harness calls are instantiated.  weights(vf): the field given at construction, else the mesh's linear
// weights (linear.H:88-97).
struct surfaceInterpolation {
    static int debug;
};
template <class Type> class surfaceInterpolationScheme : public refCount
{
    const fvMesh *mesh_ = nullptr;
    const surfaceScalarField *w_ = nullptr;

public:
    struct ConstructorTableStub {
        struct iterator {
            bool operator==(const iterator &) const { return true; }
            struct Maker {
                template <class... A> tmp<surfaceInterpolationScheme<Type>> operator()(const A &...) const
                {
                    throw std::runtime_error("no run-time selection in the harness");
                }
            };
            Maker operator()() const { return Maker(); }
        };
        iterator find(const word &) { return iterator(); }
        iterator end() { return iterator(); }
        word sortedToc() const { return word(); }
    };
    typedef ConstructorTableStub MeshConstructorTable;
    typedef ConstructorTableStub MeshFluxConstructorTable;
    static MeshConstructorTable *MeshConstructorTablePtr_;
    static MeshFluxConstructorTable *MeshFluxConstructorTablePtr_;
    surfaceInterpolationScheme(const fvMesh &mesh) : mesh_(&mesh) {}
    surfaceInterpolationScheme(const surfaceScalarField &w) : w_(&w) {}
    static tmp<surfaceInterpolationScheme<Type>> New(const fvMesh &mesh, Istream &schemeData);
    static tmp<surfaceInterpolationScheme<Type>> New(const fvMesh &mesh, const surfaceScalarField &faceFlux, Istream &schemeData);
    ~surfaceInterpolationScheme();
    const fvMesh &mesh() const { return *mesh_; }
    tmp<surfaceScalarField> weights(const GeometricField<Type, fvPatchField, volMesh> &) const
    {
        return tmp<surfaceScalarField>(new surfaceScalarField(w_ ? *w_ : mesh_->weights()));
    }
    bool corrected() const { return false; }
    template <class F> tmp<GeometricField<Type, fvsPatchField, surfaceMesh>> correction(const F &) const
    {
        throw std::runtime_error("not used by the harness");
    }
    static tmp<GeometricField<Type, fvsPatchField, surfaceMesh>> interpolate(const GeometricField<Type, fvPatchField, volMesh> &,
                                                                             const tmp<surfaceScalarField> &,
                                                                             const tmp<surfaceScalarField> &);
    static tmp<GeometricField<Type, fvsPatchField, surfaceMesh>> interpolate(const GeometricField<Type, fvPatchField, volMesh> &,
                                                                             const tmp<surfaceScalarField> &);
    tmp<GeometricField<Type, fvsPatchField, surfaceMesh>> interpolate(const GeometricField<Type, fvPatchField, volMesh> &) const;
    tmp<GeometricField<Type, fvsPatchField, surfaceMesh>> interpolate(const tmp<GeometricField<Type, fvPatchField, volMesh>> &) const;
};
template <class Type> class snGradScheme : public refCount
{
public:
    bool corrected() const { return false; }
    template <class F> tmp<surfaceScalarField> deltaCoeffs(const F &) const { throw std::runtime_error("not used by the harness"); }
    template <class F> tmp<GeometricField<Type, fvsPatchField, surfaceMesh>> snGrad(const F &) const
    {
        throw std::runtime_error("not used by the harness");
    }
    template <class F> tmp<GeometricField<Type, fvsPatchField, surfaceMesh>> correction(const F &) const
    {
        throw std::runtime_error("not used by the harness");
    }
};
namespace fvc
{
template <class T> tmp<GeometricField<T, fvPatchField, volMesh>> surfaceIntegrate(const tmp<GeometricField<T, fvsPatchField, surfaceMesh>> &)
{
    throw std::runtime_error("not used by the harness");
}
template <class F> F div(const F &) { throw std::runtime_error("not used by the harness"); }
template <class F> F grad(const F &) { throw std::runtime_error("not used by the harness"); }
template <class F> F interpolate(const F &) { throw std::runtime_error("not used by the harness"); }
} // namespace fvc
#define TypeName(name)                              \
    static const char *typeName_() { return name; } \
    static const ::Foam::word typeName;             \
    static int debug
namespace fv
{
template <class Type> class convectionScheme : public refCount // convectionScheme.H
{
    const fvMesh &mesh_;

public:
    convectionScheme(const fvMesh &m, const surfaceScalarField &) : mesh_(m) {}
    virtual ~convectionScheme() {}
    const fvMesh &mesh() const { return mesh_; }
};
template <class Type> class gaussConvectionScheme : public fv::convectionScheme<Type> // gaussConvectionScheme.H:55-170
{
    tmp<surfaceInterpolationScheme<Type>> tinterpScheme_;

public:
    gaussConvectionScheme(const fvMesh &mesh, const surfaceScalarField &faceFlux, const tmp<surfaceInterpolationScheme<Type>> &scheme)
        : convectionScheme<Type>(mesh, faceFlux), tinterpScheme_(scheme)
    {
    }
    const surfaceInterpolationScheme<Type> &interpScheme() const;
    tmp<GeometricField<Type, fvsPatchField, surfaceMesh>> interpolate(const surfaceScalarField &,
                                                                      const GeometricField<Type, fvPatchField, volMesh> &) const;
    tmp<GeometricField<Type, fvsPatchField, surfaceMesh>> flux(const surfaceScalarField &,
                                                               const GeometricField<Type, fvPatchField, volMesh> &) const;
    tmp<fvMatrix<Type>> fvmDiv(const surfaceScalarField &, const GeometricField<Type, fvPatchField, volMesh> &) const;
    tmp<GeometricField<Type, fvPatchField, volMesh>> fvcDiv(const surfaceScalarField &,
                                                            const GeometricField<Type, fvPatchField, volMesh> &) const;
};
template <class Type, class GType> class laplacianScheme : public refCount // laplacianScheme.H
{
    const fvMesh &mesh_;

protected:
    tmp<surfaceInterpolationScheme<GType>> tinterpGammaScheme_;
    tmp<snGradScheme<Type>> tsnGradScheme_;

public:
    laplacianScheme(const fvMesh &m) : mesh_(m) {}
    virtual ~laplacianScheme() {}
    const fvMesh &mesh() const { return mesh_; }
};
template <class Type, class GType> class gaussLaplacianScheme : public fv::laplacianScheme<Type, GType> // gaussLaplacianScheme.H:55-150
{
    tmp<GeometricField<Type, fvsPatchField, surfaceMesh>> gammaSnGradCorr(const surfaceVectorField &SfGammaCorr,
                                                                          const GeometricField<Type, fvPatchField, volMesh> &);

public:
    gaussLaplacianScheme(const fvMesh &mesh) : laplacianScheme<Type, GType>(mesh) {}
    static tmp<fvMatrix<Type>> fvmLaplacianUncorrected(const surfaceScalarField &gammaMagSf, const surfaceScalarField &deltaCoeffs,
                                                       const GeometricField<Type, fvPatchField, volMesh> &);
    tmp<GeometricField<Type, fvPatchField, volMesh>> fvcLaplacian(const GeometricField<Type, fvPatchField, volMesh> &);
    tmp<fvMatrix<Type>> fvmLaplacian(const GeometricField<GType, fvsPatchField, surfaceMesh> &,
                                     const GeometricField<Type, fvPatchField, volMesh> &);
    tmp<GeometricField<Type, fvPatchField, volMesh>> fvcLaplacian(const GeometricField<GType, fvsPatchField, surfaceMesh> &,
                                                                  const GeometricField<Type, fvPatchField, volMesh> &);
};
} // namespace fv
} // namespace Foam
#endif
