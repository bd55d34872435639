// layout_shared.cu -- the "shared-coefficient" banded layout for SYMMETRIC matrices.
//
// In LDU storage a symmetric matrix keeps one coefficient per face (upper[f] serves the
// owner row and the neighbour row; LDU/lduMatrix/lduMatrix.C:328-345).  The general banded
// layout (layout.cu) stores it twice, once per row.  Here each 64-row slice streams
//   * value slots: the owner-side coefficient of every row (slot-major, 128-bit loads),
//     then the interface coefficients, then the slice's "extras": coefficients of faces
//     whose owner row lies outside the slice but whose neighbour row is inside;
//   * neighbour slots: no coefficient, only (column, ref) 16-bit pairs, ref pointing into
//     the slice's value stream, which the warp has staged in shared memory.
// Per face that is 8 B (value) + 2 B (owner column) + 4 B (neighbour column + ref)
// + a small share of extras: the LDU-algorithmic 16 B/face instead of 20 B/face.
// Row sums keep the reference's order (owner faces, neighbour faces, interface faces).
#include <algorithm>

#include "internal.h"

int layout_build_shared(b200ldu_addr *a)
{
    if (a->sharedBuilt) return B200LDU_OK;
    a->sharedBuilt = true; // one attempt
    const int nCells = a->nCells, nFaces = a->nFaces;
    const std::vector<int> &l = a->l, &u = a->u, &perm = a->perm_h, &iperm = a->iperm_h;
    const int nPad = a->L.nPad, bandRows = a->L.bandRows;
    const int nSlices = nPad / SLICE_ROWS;
    const int nRecv = a->L.nRecv;

    std::vector<int> ownerStart((size_t)nCells + 1, 0), losortStart((size_t)nCells + 1, 0), losort(nFaces);
    for (int f = 0; f < nFaces; f++) ownerStart[l[f] + 1]++;
    for (int c = 0; c < nCells; c++) ownerStart[c + 1] += ownerStart[c];
    for (int f = 0; f < nFaces; f++) losortStart[u[f] + 1]++;
    for (int c = 0; c < nCells; c++) losortStart[c + 1] += losortStart[c];
    {
        std::vector<int> cur(losortStart.begin(), losortStart.end() - 1);
        for (int f = 0; f < nFaces; f++) losort[cur[u[f]]++] = f;
    }
    std::vector<int> pfStart((size_t)nCells + 1, 0), pfItem(std::max(nRecv, 1));
    for (int i = 0; i < nRecv; i++) pfStart[a->faceCells[i] + 1]++;
    for (int c = 0; c < nCells; c++) pfStart[c + 1] += pfStart[c];
    {
        std::vector<int> cur(pfStart.begin(), pfStart.end() - 1);
        for (int i = 0; i < nRecv; i++) pfItem[cur[a->faceCells[i]]++] = i;
    }
    // halo lists of the general layout give the 16-bit columns (same psi tile)
    std::vector<int> haloStart(a->L.nBands + 1), haloIdx;
    if (a->hostOnly) {
        haloStart = a->dbg_haloStart;
        haloIdx = a->dbg_haloIdx;
    } else {
        haloIdx.resize((size_t)std::max<long long>(a->nHaloTotal, 1));
        if (cudaMemcpy(haloStart.data(), a->d_haloStart, sizeof(int) * haloStart.size(), cudaMemcpyDeviceToHost) !=
                cudaSuccess ||
            cudaMemcpy(haloIdx.data(), a->d_haloIdx, sizeof(int) * (size_t)a->nHaloTotal, cudaMemcpyDeviceToHost) !=
                cudaSuccess) {
            b200_set_error("layout_build_shared: cannot read halo lists back");
            return B200LDU_ECUDA;
        }
    }

    // ---- pass 1: slice shapes ----
    std::vector<uint16_t> VS(nSlices, 0), WO(nSlices, 0), WN(nSlices, 0);
    std::vector<int> E(nSlices, 0);
    std::vector<long long> vStart((size_t)nSlices + 1, 0), nStart((size_t)nSlices + 1, 0);
    long long tooBig = 0;
#pragma omp parallel for schedule(static) reduction(+ : tooBig)
    for (int s = 0; s < nSlices; s++) {
        int wo = 0, wn = 0, wi = 0, e = 0;
        const int r0 = s * SLICE_ROWS;
        for (int q = 0; q < SLICE_ROWS; q++) {
            int c = iperm[r0 + q];
            if (c < 0) continue;
            wo = std::max(wo, ownerStart[c + 1] - ownerStart[c]);
            wn = std::max(wn, losortStart[c + 1] - losortStart[c]);
            wi = std::max(wi, pfStart[c + 1] - pfStart[c]);
            for (int k = losortStart[c]; k < losortStart[c + 1]; k++) {
                int ro = perm[l[losort[k]]];
                if (ro < r0 || ro >= r0 + SLICE_ROWS) e++;
            }
        }
        WO[s] = (uint16_t)wo;
        WN[s] = (uint16_t)wn;
        VS[s] = (uint16_t)(wo + wi);
        E[s] = e;
        if ((long long)(wo + wi) * SLICE_ROWS + e + 2 > 65535) tooBig++;
    }
    if (tooBig) return B200LDU_OK; // shared format not representable: the general layout stays in use
    int maxWarpDoubles = 0;
    for (int s = 0; s < nSlices; s++) {
        int nv = VS[s] * SLICE_ROWS + ((E[s] + 1 + 7) & ~7); // extras + one zero pad; 64-byte multiples (cp.async)
        vStart[s + 1] = vStart[s] + nv;
        nStart[s + 1] = nStart[s] + (long long)WN[s] * SLICE_ROWS;
        maxWarpDoubles = std::max(maxWarpDoubles, nv);
    }
    int maxVS = 0, maxWN = 0;
    for (int s = 0; s < nSlices; s++) {
        maxVS = std::max(maxVS, (int)VS[s]);
        maxWN = std::max(maxWN, (int)WN[s]);
    }
    // per-warp stream buffer: values | 16-bit columns of the value slots | neighbour entries
    const int colOff = maxWarpDoubles * 8;
    const int nbrOff = colOff + maxVS * SLICE_ROWS * 2;
    const int bufBytes = (nbrOff + maxWN * SLICE_ROWS * 4 + 15) & ~15;
    // shared memory budget: psi tile + 8 double-buffered warp buffers, >= 2 CTAs per SM
    size_t smem = sizeof(double) * (size_t)((bandRows + a->L.maxHalo + 1) & ~1) + (size_t)(ENGINE_THREADS / 32) * 2 * bufBytes;
    if (smem > 100 * 1024) return B200LDU_OK;

    const long long nV = vStart[nSlices], nN = nStart[nSlices];
    std::vector<uint16_t> colV((size_t)std::max<long long>(nV, 1), 0);
    std::vector<int> codeV((size_t)std::max<long long>(nV, 1), -1);
    std::vector<uint32_t> nbr((size_t)std::max<long long>(nN, 1), 0);

    // ---- pass 2: fill ----
#pragma omp parallel for schedule(dynamic, 64)
    for (int s = 0; s < nSlices; s++) {
        const int r0 = s * SLICE_ROWS;
        const int band = r0 / bandRows, b0 = band * bandRows, b1 = b0 + bandRows;
        const int *h0 = haloIdx.data() + haloStart[band], *h1 = haloIdx.data() + haloStart[band + 1];
        auto colOf = [&](int t) -> uint16_t {
            if (t >= b0 && t < b1) return (uint16_t)(t - b0);
            return (uint16_t)(bandRows + (int)(std::lower_bound(h0, h1, t) - h0));
        };
        const long long vb = vStart[s], nb = nStart[s];
        const int vs = VS[s], wo = WO[s], wn = WN[s];
        const int extraBase = vs * SLICE_ROWS;
        int e = 0;
        const int zeroRef = extraBase + E[s]; // the padding coefficient (code -1 => 0.0)
        for (int q = 0; q < SLICE_ROWS; q++) {
            const int r = r0 + q, c = iperm[r];
            const uint16_t self = (uint16_t)(r - b0);
            for (int j = 0; j < vs; j++) colV[vb + (long long)j * SLICE_ROWS + q] = self;
            for (int j = 0; j < wn; j++) nbr[nb + (long long)j * SLICE_ROWS + q] = (uint32_t)self | ((uint32_t)zeroRef << 16);
            if (c < 0) continue;
            int j = 0;
            for (int f = ownerStart[c]; f < ownerStart[c + 1]; f++, j++) {
                colV[vb + (long long)j * SLICE_ROWS + q] = colOf(perm[u[f]]);
                codeV[vb + (long long)j * SLICE_ROWS + q] = 2 * f;
            }
            j = wo;
            for (int k = pfStart[c]; k < pfStart[c + 1]; k++, j++) {
                colV[vb + (long long)j * SLICE_ROWS + q] = colOf(nPad + pfItem[k]);
                codeV[vb + (long long)j * SLICE_ROWS + q] = -2 - pfItem[k];
            }
        }
        // neighbour slots need the owner-slot index of the face inside the owner's row
        for (int q = 0; q < SLICE_ROWS; q++) {
            const int r = r0 + q, c = iperm[r];
            if (c < 0) continue;
            int j = 0;
            for (int k = losortStart[c]; k < losortStart[c + 1]; k++, j++) {
                const int f = losort[k];
                const int o = l[f], ro = perm[o];
                int ref;
                if (ro >= r0 && ro < r0 + SLICE_ROWS) {
                    ref = (f - ownerStart[o]) * SLICE_ROWS + (ro - r0);
                } else {
                    ref = extraBase + e;
                    codeV[vb + extraBase + e] = 2 * f;
                    e++;
                }
                nbr[nb + (long long)j * SLICE_ROWS + q] = (uint32_t)colOf(ro) | ((uint32_t)ref << 16);
            }
        }
    }

    a->sh_nV = nV;
    a->sh_nN = nN;
    a->sh_warpDoubles = maxWarpDoubles;
    if (a->hostOnly) {
        a->dbg_shVStart.swap(vStart);
        a->dbg_shNStart.swap(nStart);
        a->dbg_shVS.swap(VS);
        a->dbg_shWO.swap(WO);
        a->dbg_shWN.swap(WN);
        a->dbg_shColV.swap(colV);
        a->dbg_shCodeV.swap(codeV);
        a->dbg_shNbr.swap(nbr);
        a->sharedOk = true;
        return B200LDU_OK;
    }
    TRY(dev_upload(&a->d_shVStart, vStart));
    TRY(dev_upload(&a->d_shNStart, nStart));
    TRY(dev_upload(&a->d_shVS, VS));
    TRY(dev_upload(&a->d_shWO, WO));
    TRY(dev_upload(&a->d_shWN, WN));
    TRY(dev_upload(&a->d_shColV, colV));
    TRY(dev_upload(&a->d_shCodeV, codeV));
    TRY(dev_upload(&a->d_shNbr, nbr));
    a->L.sh_vStart = a->d_shVStart;
    a->L.sh_nStart = a->d_shNStart;
    a->L.sh_VS = a->d_shVS;
    a->L.sh_WO = a->d_shWO;
    a->L.sh_WN = a->d_shWN;
    a->L.sh_colV = a->d_shColV;
    a->L.sh_nbr = a->d_shNbr;
    a->L.sh_warpDoubles = maxWarpDoubles;
    a->L.sh_bufBytes = bufBytes;
    a->L.sh_colOff = colOff;
    a->L.sh_nbrOff = nbrOff;
    a->sharedOk = true;
    return B200LDU_OK;
}
