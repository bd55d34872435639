/*
 * harness_lduaddr.cpp -- runs the REFERENCE'S OWN lduAddressing.C (derived addressing arrays) on the
 * CPU.  TEST INFRASTRUCTURE ONLY.  Included by path from /root/reference:
 *   LDU/lduAddressing/lduAddressing.C   calcLosort :169-199, calcOwnerStart :202-267, calcLosortStart
 *       :270-344, ownerSortAddr :373-400, calcPatchSort / calcPatchSortStart :38-167, band :453-498
 * against oracle/ref_harness/shim_lduaddr/.
 */
#include "lduAddressing.H" /* shim */

#include "lduAddressing.C" /* reference */

using namespace Foam;

namespace
{
class Addr : public lduAddressing
{
    labelgpuList l_, u_;
    std::vector<labelgpuList> patches_;

public:
    Addr(int n, int nF, const int *l, const int *u, int nPatches, const int *patchStart, const int *faceCells)
        : lduAddressing(n), l_(labelList(l, nF)), u_(labelList(u, nF))
    {
        for (int p = 0; p < nPatches; p++)
            patches_.push_back(labelgpuList(labelList(faceCells + patchStart[p], patchStart[p + 1] - patchStart[p])));
    }
    virtual const labelgpuList &lowerAddr() const { return l_; }
    virtual const labelgpuList &upperAddr() const { return u_; }
    virtual const labelgpuList &patchAddr(const label p) const { return patches_[(size_t)p]; }
    virtual label nPatches() const { return (label)patches_.size(); }
    virtual bool patchAvailable(const label) const { return true; }
};
void put(const labelgpuList &src, int *dst)
{
    for (label i = 0; i < src.size(); i++) dst[i] = src[i];
}
} // namespace

extern "C" {
/* outputs: ownerStart[n+1], losortStart[n+1], losort[nF], ownerSort[nF]; per patch p (concatenated at
 * patchStart[p]): sortAddr (faces of the patch ordered by cell), sortCells (unique cells, -1 padded),
 * sortStart (first sorted face of every unique cell, then the patch size; -1 padded); nUnique[p];
 * band[2] = bandwidth, profile. */
int ref_ldu_addressing(int n, int nF, const int *l, const int *u, int nPatches, const int *patchStart, const int *faceCells,
                       int *ownerStart, int *losortStart, int *losort, int *ownerSort, int *patchSortAddr,
                       int *patchSortCells, int *patchSortStart, int *nUnique, double *band)
{
    try {
        Addr a(n, nF, l, u, nPatches, patchStart, faceCells);
        put(a.ownerStartAddr(), ownerStart);
        put(a.losortStartAddr(), losortStart);
        put(a.losortAddr(), losort);
        put(a.ownerSortAddr(), ownerSort);
        for (int p = 0; p < nPatches; p++) {
            const int s = patchStart[p], np = patchStart[p + 1] - s;
            for (int i = 0; i < np; i++) patchSortCells[s + i] = patchSortStart[s + i] = -1;
            put(a.patchSortAddr(p), patchSortAddr + s);
            const labelgpuList &cells = a.patchSortCells(p);
            put(cells, patchSortCells + s);
            nUnique[p] = cells.size();
            const labelgpuList &st = a.patchSortStartAddr(p);
            for (label i = 0; i < cells.size() && i < np; i++) patchSortStart[s + i] = st[i];
        }
        Tuple2<label, scalar> b = a.band();
        band[0] = b.first();
        band[1] = b.second();
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
}
