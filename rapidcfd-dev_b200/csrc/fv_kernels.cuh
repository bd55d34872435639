// fv_kernels.cuh -- the cell-parallel face-sum kernels of csrc/fv.cu (device code only, free of launch syntax so that
// tests/host_kernels/ can run it on the host against the oracle).
//
// One thread per cell walks its owner faces, its neighbour faces (losort) and its boundary faces in the reference's order
// (fvcSurfaceIntegrate.C:136-205, gaussGrad.C:34-125, lduMatrixOperations.C:59-80), every product and sum rounded on its own.
//
// Load scheduling (round 2, after profiles/r02_ncu_fv_kernels_summary.csv showed these kernels latency-bound at ~3.3 TB/s with
// DRAM traffic already minimal): nvcc does not hoist a load above a loop of unknown trip count, so the straightforward
// "owner loop, neighbour loop, boundary loop, divide" issued its global loads as a chain six to eight deep (ownerStart ->
// owner values -> losortStart -> losort -> neighbour values -> bStart -> bFaces -> boundary values -> V).  Here everything a
// thread can know up front is loaded first (the three row ranges, V, the cell's own value), then the first batch of owner
// values TOGETHER with the first batch of losort indices and the first boundary face index, then the values those indices
// point at: three levels.  Cells with more than FV_BATCH faces on a side continue with plain loops, in the same order.
// Measured at 256^3 (profiles/r02_kernel_table_n256.txt against ..._before_hoisting.txt): surfaceIntegrate scalar 258 -> 231 us,
// gaussGrad scalar 562 -> 533 us, negSumDiag fills 385 -> 355 / 468 -> 443 us, fvMatrix::H 422 -> 359 / 617 -> 518 us, relax
// 406 -> 379 / 445 -> 427 us.  Where the scheme needs ~64 registers it loses more occupancy than the shorter chain gains
// (surfaceIntegrate of a vector field, the fused interpolate + gradient, the MULES sweeps): those keep their plain loops.
#ifndef B200LDU_FV_KERNELS_CUH
#define B200LDU_FV_KERNELS_CUH
#include <cstddef>

namespace fvk
{
namespace
{
constexpr int FV_BATCH = 3;

template <int NC>
__global__ void surface_integrate_kernel(int nCells, const int *__restrict__ ownerStart,
                                         const int *__restrict__ losortStart, const int *__restrict__ losort,
                                         const int *__restrict__ bStart, const int *__restrict__ bFaces,
                                         const double *__restrict__ ssf, const double *__restrict__ bssf,
                                         const double *__restrict__ V, double *__restrict__ out, int divideByV,
                                         int neiSign)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    // level 1: the row ranges and the volume
    const int o0 = ownerStart[c], o1 = ownerStart[c + 1], n0 = losortStart[c], n1 = losortStart[c + 1];
    int b0 = 0, b1 = 0;
    if (bStart) b0 = bStart[c], b1 = bStart[c + 1];
    const double vol = divideByV ? V[c] : 1.0;
    double acc[NC];
#pragma unroll
    for (int k = 0; k < NC; k++) acc[k] = 0.0;
    if constexpr (NC == 1) {
    // level 2: first batch of owner values, of losort indices, the first boundary face
    double ov[FV_BATCH][NC];
    int fi[FV_BATCH], bf0 = 0;
#pragma unroll
    for (int b = 0; b < FV_BATCH; b++)
        if (o0 + b < o1)
#pragma unroll
            for (int k = 0; k < NC; k++) ov[b][k] = ssf[(size_t)(o0 + b) * NC + k];
#pragma unroll
    for (int b = 0; b < FV_BATCH; b++)
        if (n0 + b < n1) fi[b] = losort[n0 + b];
    if (b0 < b1) bf0 = bFaces[b0];
    // level 3: the values behind those indices
    double nv[FV_BATCH][NC], bv0[NC];
#pragma unroll
    for (int b = 0; b < FV_BATCH; b++)
        if (n0 + b < n1)
#pragma unroll
            for (int k = 0; k < NC; k++) nv[b][k] = ssf[(size_t)fi[b] * NC + k];
    if (b0 < b1)
#pragma unroll
        for (int k = 0; k < NC; k++) bv0[k] = bssf[(size_t)bf0 * NC + k];
    // sums in face order
#pragma unroll
    for (int b = 0; b < FV_BATCH; b++)
        if (o0 + b < o1)
#pragma unroll
            for (int k = 0; k < NC; k++) acc[k] = __dadd_rn(acc[k], ov[b][k]);
    for (int f = o0 + FV_BATCH; f < o1; f++)
#pragma unroll
        for (int k = 0; k < NC; k++) acc[k] = __dadd_rn(acc[k], ssf[(size_t)f * NC + k]);
#pragma unroll
    for (int b = 0; b < FV_BATCH; b++)
        if (n0 + b < n1)
#pragma unroll
            for (int k = 0; k < NC; k++) acc[k] = neiSign < 0 ? __dsub_rn(acc[k], nv[b][k]) : __dadd_rn(acc[k], nv[b][k]);
    for (int j = n0 + FV_BATCH; j < n1; j++) {
        const int f = losort[j];
#pragma unroll
        for (int k = 0; k < NC; k++) {
            const double v = ssf[(size_t)f * NC + k];
            acc[k] = neiSign < 0 ? __dsub_rn(acc[k], v) : __dadd_rn(acc[k], v);
        }
    }
    if (b0 < b1) {
#pragma unroll
        for (int k = 0; k < NC; k++) acc[k] = __dadd_rn(acc[k], bv0[k]);
        for (int j = b0 + 1; j < b1; j++) {
            const int bf = bFaces[j];
#pragma unroll
            for (int k = 0; k < NC; k++) acc[k] = __dadd_rn(acc[k], bssf[(size_t)bf * NC + k]);
        }
    }
    } else {
        // vector fields: holding two batches of three-component values costs the occupancy more than the shorter chain gains
        // (measured: 484 us against 453 us at 256^3), so only the row ranges are loaded ahead
        for (int f = o0; f < o1; f++)
#pragma unroll
            for (int k = 0; k < NC; k++) acc[k] = __dadd_rn(acc[k], ssf[(size_t)f * NC + k]);
        for (int j = n0; j < n1; j++) {
            const int f = losort[j];
#pragma unroll
            for (int k = 0; k < NC; k++) {
                const double v = ssf[(size_t)f * NC + k];
                acc[k] = neiSign < 0 ? __dsub_rn(acc[k], v) : __dadd_rn(acc[k], v);
            }
        }
        for (int j = b0; j < b1; j++) {
            const int bf = bFaces[j];
#pragma unroll
            for (int k = 0; k < NC; k++) acc[k] = __dadd_rn(acc[k], bssf[(size_t)bf * NC + k]);
        }
    }
#pragma unroll
    for (int k = 0; k < NC; k++) out[(size_t)c * NC + k] = divideByV ? __ddiv_rn(acc[k], vol) : acc[k];
}

// NC = 1: vector result; NC = 3: tensor result T[i][j] = Sf[i]*ssf[j]
template <int NC>
__global__ void gauss_grad_kernel(int nCells, const int *__restrict__ ownerStart,
                                  const int *__restrict__ losortStart, const int *__restrict__ losort,
                                  const int *__restrict__ bStart, const int *__restrict__ bFaces,
                                  const double *__restrict__ Sf, const double *__restrict__ ssf,
                                  const double *__restrict__ bSf, const double *__restrict__ bssf,
                                  const double *__restrict__ V, double *__restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    const int o0 = ownerStart[c], o1 = ownerStart[c + 1], n0 = losortStart[c], n1 = losortStart[c + 1];
    int b0 = 0, b1 = 0;
    if (bStart) b0 = bStart[c], b1 = bStart[c + 1];
    const double vol = V[c];
    int fi[FV_BATCH], bf0 = 0;
#pragma unroll
    for (int b = 0; b < FV_BATCH; b++)
        if (n0 + b < n1) fi[b] = losort[n0 + b];
    if (b0 < b1) bf0 = bFaces[b0];
    double acc[3 * NC];
#pragma unroll
    for (int k = 0; k < 3 * NC; k++) acc[k] = 0.0;
    { // owner faces: first batch with all loads ahead of the products, the rest one by one
        double sv[FV_BATCH][3], fv[FV_BATCH][NC];
#pragma unroll
        for (int b = 0; b < FV_BATCH; b++)
            if (o0 + b < o1) {
                const size_t f = (size_t)(o0 + b);
                sv[b][0] = Sf[f * 3], sv[b][1] = Sf[f * 3 + 1], sv[b][2] = Sf[f * 3 + 2];
#pragma unroll
                for (int j = 0; j < NC; j++) fv[b][j] = ssf[f * NC + j];
            }
#pragma unroll
        for (int b = 0; b < FV_BATCH; b++)
            if (o0 + b < o1)
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int j = 0; j < NC; j++) acc[i * NC + j] = __dadd_rn(acc[i * NC + j], __dmul_rn(sv[b][i], fv[b][j]));
        for (int f = o0 + FV_BATCH; f < o1; f++) {
            const double s[3] = {Sf[(size_t)f * 3], Sf[(size_t)f * 3 + 1], Sf[(size_t)f * 3 + 2]};
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < NC; j++) acc[i * NC + j] = __dadd_rn(acc[i * NC + j], __dmul_rn(s[i], ssf[(size_t)f * NC + j]));
        }
    }
    { // neighbour faces
        double sv[FV_BATCH][3], fv[FV_BATCH][NC];
#pragma unroll
        for (int b = 0; b < FV_BATCH; b++)
            if (n0 + b < n1) {
                const size_t f = (size_t)fi[b];
                sv[b][0] = Sf[f * 3], sv[b][1] = Sf[f * 3 + 1], sv[b][2] = Sf[f * 3 + 2];
#pragma unroll
                for (int j = 0; j < NC; j++) fv[b][j] = ssf[f * NC + j];
            }
#pragma unroll
        for (int b = 0; b < FV_BATCH; b++)
            if (n0 + b < n1)
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int j = 0; j < NC; j++) acc[i * NC + j] = __dsub_rn(acc[i * NC + j], __dmul_rn(sv[b][i], fv[b][j]));
        for (int q = n0 + FV_BATCH; q < n1; q++) {
            const size_t f = (size_t)losort[q];
            const double s[3] = {Sf[f * 3], Sf[f * 3 + 1], Sf[f * 3 + 2]};
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < NC; j++) acc[i * NC + j] = __dsub_rn(acc[i * NC + j], __dmul_rn(s[i], ssf[f * NC + j]));
        }
    }
    for (int q = b0; q < b1; q++) {
        const int bf = q == b0 ? bf0 : bFaces[q];
        const double s[3] = {bSf[(size_t)bf * 3], bSf[(size_t)bf * 3 + 1], bSf[(size_t)bf * 3 + 2]};
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < NC; j++)
                acc[i * NC + j] = __dadd_rn(acc[i * NC + j], __dmul_rn(s[i], bssf[(size_t)bf * NC + j]));
    }
#pragma unroll
    for (int k = 0; k < 3 * NC; k++) out[(size_t)c * 3 * NC + k] = __ddiv_rn(acc[k], vol);
}

// diag[c] = 0 - sum_{own} lower[f] - sum_{nei} upper[f]   (negSumDiag, lduMatrixOperations.C:59-80)
__global__ void neg_sum_diag_kernel(int nCells, const int *__restrict__ ownerStart,
                                    const int *__restrict__ losortStart, const int *__restrict__ losort,
                                    const double *__restrict__ upper, const double *__restrict__ lower,
                                    double *__restrict__ diag)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    const int o0 = ownerStart[c], o1 = ownerStart[c + 1], n0 = losortStart[c], n1 = losortStart[c + 1];
    double ov[FV_BATCH], nv[FV_BATCH];
    int fi[FV_BATCH];
#pragma unroll
    for (int b = 0; b < FV_BATCH; b++)
        if (o0 + b < o1) ov[b] = lower[o0 + b];
#pragma unroll
    for (int b = 0; b < FV_BATCH; b++)
        if (n0 + b < n1) fi[b] = losort[n0 + b];
#pragma unroll
    for (int b = 0; b < FV_BATCH; b++)
        if (n0 + b < n1) nv[b] = upper[fi[b]];
    double acc = 0.0;
#pragma unroll
    for (int b = 0; b < FV_BATCH; b++)
        if (o0 + b < o1) acc = __dsub_rn(acc, ov[b]);
    for (int f = o0 + FV_BATCH; f < o1; f++) acc = __dsub_rn(acc, lower[f]);
#pragma unroll
    for (int b = 0; b < FV_BATCH; b++)
        if (n0 + b < n1) acc = __dsub_rn(acc, nv[b]);
    for (int j = n0 + FV_BATCH; j < n1; j++) acc = __dsub_rn(acc, upper[losort[j]]);
    diag[c] = acc;
}

// linear face value as interpolate(vf) forms it: w*(own - nei) + nei (surfaceInterpolationScheme.C:272-351), each operation
// rounded on its own
__device__ __forceinline__ double lin_face(double w, double own, double nei)
{
    return __dadd_rn(__dmul_rn(w, __dsub_rn(own, nei)), nei);
}

// gaussGrad::calcGrad = gradf(interpolate(vsf)) (gaussGrad.C:256-271 with the linear scheme) without the F-sized face field:
// the face value w*(psi[own] - psi[nei]) + psi[nei] is formed on the fly with the same rounded subtract, product and add as
// the unfused pipeline, so the result equals gauss_grad(interpolate_linear(...)) bit for bit.
template <int NC>
__global__ void grad_linear_kernel(int nCells, const int *__restrict__ ownerStart, const int *__restrict__ upper,
                                   const int *__restrict__ losortStart, const int *__restrict__ losort,
                                   const int *__restrict__ lower, const int *__restrict__ bStart,
                                   const int *__restrict__ bFaces, const double *__restrict__ Sf,
                                   const double *__restrict__ w, const double *__restrict__ vf,
                                   const double *__restrict__ bSf, const double *__restrict__ bvf,
                                   const double *__restrict__ V, double *__restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    // plain loops: prefetching the first batch of indices (the scheme of the kernels above) raised this kernel to 64 registers
    // and made it slower (791 us against 664 us at 256^3, scalar field)
    double mine[NC];
#pragma unroll
    for (int j = 0; j < NC; j++) mine[j] = vf[(size_t)c * NC + j];
    double acc[3 * NC];
#pragma unroll
    for (int k = 0; k < 3 * NC; k++) acc[k] = 0.0;
    for (int f = ownerStart[c]; f < ownerStart[c + 1]; f++) {
        const int n = upper[f];
        const double ww = w[f];
        const double s[3] = {Sf[(size_t)f * 3], Sf[(size_t)f * 3 + 1], Sf[(size_t)f * 3 + 2]};
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const double fv = lin_face(ww, mine[j], vf[(size_t)n * NC + j]);
#pragma unroll
            for (int i = 0; i < 3; i++) acc[i * NC + j] = __dadd_rn(acc[i * NC + j], __dmul_rn(s[i], fv));
        }
    }
    for (int q = losortStart[c]; q < losortStart[c + 1]; q++) {
        const int f = losort[q], o = lower[f];
        const double ww = w[f];
        const double s[3] = {Sf[(size_t)f * 3], Sf[(size_t)f * 3 + 1], Sf[(size_t)f * 3 + 2]};
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const double fv = lin_face(ww, vf[(size_t)o * NC + j], mine[j]);
#pragma unroll
            for (int i = 0; i < 3; i++) acc[i * NC + j] = __dsub_rn(acc[i * NC + j], __dmul_rn(s[i], fv));
        }
    }
    if (bStart)
        for (int q = bStart[c]; q < bStart[c + 1]; q++) {
            const int bf = bFaces[q];
            const double s[3] = {bSf[(size_t)bf * 3], bSf[(size_t)bf * 3 + 1], bSf[(size_t)bf * 3 + 2]};
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < NC; j++)
                    acc[i * NC + j] = __dadd_rn(acc[i * NC + j], __dmul_rn(s[i], bvf[(size_t)bf * NC + j]));
        }
    const double v = V[c];
#pragma unroll
    for (int k = 0; k < 3 * NC; k++) out[(size_t)c * 3 * NC + k] = __ddiv_rn(acc[k], v);
}
} // namespace
} // namespace fvk
#endif
