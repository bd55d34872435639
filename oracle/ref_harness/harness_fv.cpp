/*
 * harness_fv.cpp -- runs the REFERENCE'S OWN fvc::surfaceIntegrate on the CPU (scalar fields).  TEST
 * INFRASTRUCTURE ONLY.  Included by path from /root/reference:
 *   FV/finiteVolume/fvc/fvcSurfaceIntegrate.C   surfaceIntegrateFunctor, surfaceIntegratePatchFunctor,
 *                                               fvc::surfaceIntegrate(ivf, ssf) :41-205, surfaceSum :261-352
 *   FV/finiteVolume/gradSchemes/gaussGrad/gaussGrad.C   gaussGradFunctor, gaussGradPatchFunctor,
 *                                               fv::gaussGrad<scalar>::gradf :34-243
 * against oracle/ref_harness/shim_fv/.  All boundary faces are handed over as ONE patch: the reference adds
 * patch by patch and, inside a patch, cell by cell in ascending patch-face order -- the same order as one
 * concatenated patch sorted stably by cell.
 */
#include "fv_shim.h"

#include <algorithm>

#include "fvcSurfaceIntegrate.C" /* reference */
#include "gaussGrad.C"           /* reference */

int Foam::lduMatrixSolutionCache::favourSpeed = 0;

using namespace Foam;

namespace
{
struct MeshCase { // one boundary patch holding every boundary face, with its sort addressing
    fvMesh mesh;
    std::vector<label> order, cells, start;
    MeshCase(int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart, const int *losort,
             int nB, const int *bFaceCells, const double *V)
    {
        mesh.addr_.nCells_ = n;
        mesh.addr_.lower_.view(l, nF);
        mesh.addr_.upper_.view(u, nF);
        mesh.addr_.ownerStart_.view(ownerStart, n + 1);
        mesh.addr_.losortStart_.view(losortStart, n + 1);
        mesh.addr_.losort_.view(losort, nF);
        // per-patch sort addressing of the single boundary patch (lduAddressing.C:38-130)
        order.resize(nB);
        for (int i = 0; i < nB; i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](label a, label b) { return bFaceCells[a] < bFaceCells[b]; });
        for (int k = 0; k < nB; k++) {
            const label c = bFaceCells[order[k]];
            if (cells.empty() || cells.back() != c) {
                cells.push_back(c);
                start.push_back(k);
            }
        }
        start.push_back(nB);
        mesh.addr_.patchCells_.view(cells.data(), (label)cells.size());
        mesh.addr_.patchSort_.view(order.data(), nB);
        mesh.addr_.patchSortStart_.view(start.data(), (label)start.size());
        mesh.nBoundaryPatches_ = nB ? 1 : 0;
        mesh.V_.view(V, n);
    }
};
} // namespace

extern "C" {
/* integrate != 0: out[c] = (sum_owner ssf - sum_neighbour ssf + sum_boundary bssf) / V[c]   (surfaceIntegrate)
 * integrate == 0: out[c] =  sum_owner ssf + sum_neighbour ssf + sum_boundary bssf           (surfaceSum :261-352) */
int ref_surface_integrate(int integrate, int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart,
                          const int *losort, const double *ssf, int nB, const int *bFaceCells, const double *bssf,
                          const double *V, double *out)
{
    MeshCase mc(n, nF, l, u, ownerStart, losortStart, losort, nB, bFaceCells, V);
    fvMesh &mesh = mc.mesh;

    GeometricField<scalar, fvsPatchField, surfaceMesh> sf;
    sf.mesh_ = &mesh;
    sf.internal_.view(ssf, nF);
    if (nB) {
        sf.boundary_.resize(1);
        sf.boundary_[0].view(bssf, nB);
    }
    if (integrate) {
        scalargpuField ivf(out, n);
        fvc::surfaceIntegrate(ivf, sf);
    } else {
        tmp<GeometricField<scalar, fvPatchField, volMesh>> t = fvc::surfaceSum(sf);
        std::copy(t().getField().data(), t().getField().data() + n, out);
    }
    return 0;
}

/* out[c] = (sum_owner Sf*ssf - sum_neighbour Sf*ssf + sum_boundary bSf*bssf) / V[c]; Sf, bSf, out: 3 per entry */
int ref_gauss_gradf(int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart,
                    const int *losort, const double *Sf, const double *ssf, int nB, const int *bFaceCells,
                    const double *bSf, const double *bssf, const double *V, double *out)
{
    MeshCase mc(n, nF, l, u, ownerStart, losortStart, losort, nB, bFaceCells, V);
    fvMesh &mesh = mc.mesh;
    GeometricField<vector, fvsPatchField, surfaceMesh> area;
    area.mesh_ = &mesh;
    area.internal_.view(reinterpret_cast<const vector *>(Sf), nF);
    GeometricField<scalar, fvsPatchField, surfaceMesh> sf;
    sf.mesh_ = &mesh;
    sf.internal_.view(ssf, nF);
    if (nB) {
        area.boundary_.resize(1);
        area.boundary_[0].view(reinterpret_cast<const vector *>(bSf), nB);
        sf.boundary_.resize(1);
        sf.boundary_[0].view(bssf, nB);
    }
    mesh.Sf_ = &area;
    tmp<GeometricField<vector, fvPatchField, volMesh>> t = fv::gaussGrad<scalar>::gradf(sf, word("grad"));
    const vector *g = t().getField().data();
    for (int c = 0; c < n; c++)
        for (int k = 0; k < 3; k++) out[3 * c + k] = g[c].v_[k];
    return 0;
}

/* the same two face sums for a VECTOR surface field (ssf, bssf: 3 per face): surfaceIntegrate / surfaceSum -> out [n*3];
 * gaussGrad<vector>::gradf -> the tensor field, out [n*9], T_ij = d_i u_j */
int ref_surface_integrate_vec(int integrate, int n, int nF, const int *l, const int *u, const int *ownerStart,
                              const int *losortStart, const int *losort, const double *ssf, int nB, const int *bFaceCells,
                              const double *bssf, const double *V, double *out)
{
    MeshCase mc(n, nF, l, u, ownerStart, losortStart, losort, nB, bFaceCells, V);
    fvMesh &mesh = mc.mesh;
    GeometricField<vector, fvsPatchField, surfaceMesh> sf;
    sf.mesh_ = &mesh;
    sf.internal_.view(reinterpret_cast<const vector *>(ssf), nF);
    if (nB) {
        sf.boundary_.resize(1);
        sf.boundary_[0].view(reinterpret_cast<const vector *>(bssf), nB);
    }
    if (integrate) {
        gpuField<vector> ivf(reinterpret_cast<vector *>(out), n);
        fvc::surfaceIntegrate(ivf, sf);
    } else {
        tmp<GeometricField<vector, fvPatchField, volMesh>> t = fvc::surfaceSum(sf);
        std::copy(reinterpret_cast<const double *>(t().getField().data()),
                  reinterpret_cast<const double *>(t().getField().data()) + 3 * (size_t)n, out);
    }
    return 0;
}

int ref_gauss_gradf_vec(int n, int nF, const int *l, const int *u, const int *ownerStart, const int *losortStart,
                        const int *losort, const double *Sf, const double *ssf, int nB, const int *bFaceCells,
                        const double *bSf, const double *bssf, const double *V, double *out)
{
    MeshCase mc(n, nF, l, u, ownerStart, losortStart, losort, nB, bFaceCells, V);
    fvMesh &mesh = mc.mesh;
    GeometricField<vector, fvsPatchField, surfaceMesh> area, sf;
    area.mesh_ = sf.mesh_ = &mesh;
    area.internal_.view(reinterpret_cast<const vector *>(Sf), nF);
    sf.internal_.view(reinterpret_cast<const vector *>(ssf), nF);
    if (nB) {
        area.boundary_.resize(1);
        area.boundary_[0].view(reinterpret_cast<const vector *>(bSf), nB);
        sf.boundary_.resize(1);
        sf.boundary_[0].view(reinterpret_cast<const vector *>(bssf), nB);
    }
    mesh.Sf_ = &area;
    tmp<GeometricField<tensor, fvPatchField, volMesh>> t = fv::gaussGrad<vector>::gradf(sf, word("grad"));
    const tensor *g = t().getField().data();
    for (int c = 0; c < n; c++)
        for (int k = 0; k < 9; k++) out[9 * c + k] = g[c].v_[k];
    return 0;
}
}
