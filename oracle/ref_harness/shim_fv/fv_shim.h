/*
 * fv_shim.h -- the finite-volume types the reference's fvcSurfaceIntegrate.C is written against, reduced
 * to what its core template `fvc::surfaceIntegrate(gpuField<Type>&, const surfaceField&)` touches
 * (fvcSurfaceIntegrate.C:136-205): an fvMesh with the LDU addressing, the boundary list and the cell
 * volumes, a surface field with an internal and per-patch values.  The other templates of that file
 * (the tmp-returning wrappers, surfaceSum) only have to parse and are never instantiated.  TEST
 * INFRASTRUCTURE ONLY; scalar fields.
 */
#ifndef FV_SHIM_H
#define FV_SHIM_H
#include "../shim/foam_shim.h"

#include <string>

namespace Foam
{
class word : public std::string
{
public:
    word() {}
    word(const char *s) : std::string(s) {}
    word(const std::string &s) : std::string(s) {}
};
struct IOobject {
    enum readOption { NO_READ };
    enum writeOption { NO_WRITE };
    template <class... A> IOobject(const A &...) {}
};
// Vector<scalar> reduced to what gaussGrad.C:34-125 uses.  The component-wise products follow
// VectorSpaceI.H:604-630 / VectorSpaceM.H:36-43 (vs[i] = vs1[i]*s, vs[i] = vs1[i]/s) and :213-262 (+=, -=).
struct vector {
    scalar v_[3];
    vector() {}
    vector(scalar x, scalar y, scalar z) : v_{x, y, z} {}
    void operator+=(const vector &o)
    {
        for (int i = 0; i < 3; i++) v_[i] += o.v_[i];
    }
    void operator-=(const vector &o)
    {
        for (int i = 0; i < 3; i++) v_[i] -= o.v_[i];
    }
};
inline vector operator*(const vector &a, scalar s) { return vector(a.v_[0] * s, a.v_[1] * s, a.v_[2] * s); }
inline vector operator/(const vector &a, scalar s) { return vector(a.v_[0] / s, a.v_[1] / s, a.v_[2] / s); }
template <> struct pTraits<vector> {
    static const vector zero;
};
inline const vector pTraits<vector>::zero(0, 0, 0);
typedef gpuField<vector> vectorgpuField;
template <class A, class B> struct outerProduct;
template <> struct outerProduct<vector, scalar> {
    typedef vector type;
};
// Tensor<scalar> for grad(vector): T_ij = a_i*b_j (TensorI.H:439-448), component-wise +=, -=, / (VectorSpaceI.H)
struct tensor {
    scalar v_[9];
    tensor() {}
    void operator+=(const tensor &o)
    {
        for (int i = 0; i < 9; i++) v_[i] += o.v_[i];
    }
    void operator-=(const tensor &o)
    {
        for (int i = 0; i < 9; i++) v_[i] -= o.v_[i];
    }
};
inline tensor operator*(const vector &a, const vector &b)
{
    tensor t;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t.v_[3 * i + j] = a.v_[i] * b.v_[j];
    return t;
}
inline tensor operator/(const tensor &a, scalar s)
{
    tensor t;
    for (int i = 0; i < 9; i++) t.v_[i] = a.v_[i] / s;
    return t;
}
template <> struct pTraits<tensor> {
    static const tensor zero;
};
inline const tensor pTraits<tensor>::zero = [] {
    tensor t;
    for (int i = 0; i < 9; i++) t.v_[i] = 0;
    return t;
}();
template <> struct outerProduct<vector, vector> {
    typedef tensor type;
};

struct dimensionSet {
};
static const dimensionSet dimLength;
inline dimensionSet operator/(const dimensionSet &, const dimensionSet &) { return dimensionSet(); }
static const dimensionSet dimVol;
template <class Type> struct dimensioned {
    template <class... A> dimensioned(const A &...) {}
};
struct volMesh {
};
struct surfaceMesh {
};
template <class Type> struct fvPatchField {
};
template <class Type> struct zeroGradientFvPatchField {
    static const word typeName;
};
template <class Type> class fvsPatchField : public gpuField<Type>
{
public:
    using gpuField<Type>::gpuField;
};

template <class T> struct VolumesHolder { // mesh.Vsc()().getField()
    const gpuField<T> *f;
    const VolumesHolder &operator()() const { return *this; }
    const gpuField<T> &getField() const { return *f; }
};

template <class Type> class fvsPatchField;
struct surfaceMesh;
template <class Type, template <class> class PatchField, class GeoMesh> class GeometricField;

class fvMesh
{
public:
    const GeometricField<vector, fvsPatchField, surfaceMesh> *Sf_ = nullptr;
    const GeometricField<vector, fvsPatchField, surfaceMesh> &Sf() const { return *Sf_; }
    const gpuField<scalar> &V() const { return V_; } // a DimensionedField in the reference; used as `field /= V`
    lduAddressing addr_;
    label nBoundaryPatches_;
    scalargpuField V_;
    struct BoundaryList {
        label n;
        label size() const { return n; }
    };
    const lduAddressing &lduAddr() const { return addr_; }
    BoundaryList boundary() const { return BoundaryList{nBoundaryPatches_}; }
    VolumesHolder<scalar> Vsc() const { return VolumesHolder<scalar>{&V_}; }
};

template <class Type, template <class> class PatchField, class GeoMesh> class GeometricField
{
public:
    const fvMesh *mesh_;
    gpuField<Type> internal_;
    std::vector<PatchField<Type>> boundary_;
    // the constructor surfaceSum uses (fvcSurfaceIntegrate.C:272-285): a zero field on the cells of `mesh`
    GeometricField(const IOobject &, const fvMesh &mesh, const dimensioned<Type> &, const word &)
        : mesh_(&mesh), internal_(mesh.lduAddr().size(), pTraits<Type>::zero)
    {
    }
    GeometricField() : mesh_(nullptr) {}
    const fvMesh &mesh() const { return *mesh_; }
    const gpuField<Type> &getField() const { return internal_; }
    gpuField<Type> &getField() { return internal_; }
    gpuField<Type> &internalField() { return internal_; }
    const std::vector<PatchField<Type>> &boundaryField() const { return boundary_; }
    label size() const { return internal_.size(); }
    word name() const { return word("field"); }
    word instance() const { return word("0"); }
    dimensionSet dimensions() const { return dimensionSet(); }
    void correctBoundaryConditions() {}
};

inline void operator/=(gpuField<scalar> &a, const gpuField<scalar> &b)
{
    for (label i = 0; i < a.size(); i++) a.data()[i] /= b.data()[i];
}
inline void operator/=(gpuField<vector> &a, const gpuField<scalar> &b) // FieldFunctions: f[i] = f[i] / s[i]
{
    for (label i = 0; i < a.size(); i++) a.data()[i] = a.data()[i] / b.data()[i];
}
inline void operator/=(gpuField<tensor> &a, const gpuField<scalar> &b)
{
    for (label i = 0; i < a.size(); i++) a.data()[i] = a.data()[i] / b.data()[i];
}
inline word operator+(const char *a, const word &b) { return word(std::string(a) + b); }
inline word operator+(const word &a, char c) { return word(static_cast<const std::string &>(a) + c); }

template <class Type> struct surfaceInterpolationScheme {
};
namespace fv // gaussGrad.H:58-160
{
template <class Type> class gaussGrad
{
    tmp<surfaceInterpolationScheme<Type>> tinterpScheme_;

public:
    static tmp<GeometricField<typename outerProduct<vector, Type>::type, fvPatchField, volMesh>>
    gradf(const GeometricField<Type, fvsPatchField, surfaceMesh> &, const word &name);
    tmp<GeometricField<typename outerProduct<vector, Type>::type, fvPatchField, volMesh>>
    calcGrad(const GeometricField<Type, fvPatchField, volMesh> &vsf, const word &name) const; // never instantiated
    static void correctBoundaryConditions(const GeometricField<Type, fvPatchField, volMesh> &,
                                          GeometricField<typename outerProduct<vector, Type>::type, fvPatchField, volMesh> &);
};
} // namespace fv

namespace fvc // fvcSurfaceIntegrate.H
{
template <class Type>
void surfaceIntegrate(gpuField<Type> &, const GeometricField<Type, fvsPatchField, surfaceMesh> &);
template <class Type>
tmp<GeometricField<Type, fvPatchField, volMesh>> surfaceIntegrate(const GeometricField<Type, fvsPatchField, surfaceMesh> &);
template <class Type>
tmp<GeometricField<Type, fvPatchField, volMesh>>
surfaceIntegrate(const tmp<GeometricField<Type, fvsPatchField, surfaceMesh>> &);
template <class Type>
tmp<GeometricField<Type, fvPatchField, volMesh>> surfaceSum(const GeometricField<Type, fvsPatchField, surfaceMesh> &);
template <class Type>
tmp<GeometricField<Type, fvPatchField, volMesh>> surfaceSum(const tmp<GeometricField<Type, fvsPatchField, surfaceMesh>> &);
} // namespace fvc
} // namespace Foam
#endif
