/*
 * harness_gamgaddr.cpp -- runs the REFERENCE'S OWN coarse-addressing construction and level
 * combination on the CPU.  TEST INFRASTRUCTURE ONLY.  Included by path from /root/reference:
 *   GAMG/GAMGAgglomerations/GAMGAgglomeration/GAMGAgglomerateLduAddressing.C
 *       agglomerateLduAddressing :245-603, combineLevels :606-765 (+ createSort/createTarget/buildFull*)
 * against oracle/ref_harness/shim_gamgaddr/.
 */
#include "GAMGAgglomeration.H" /* shim */

#include "GAMGAgglomerateLduAddressing.C" /* reference */

namespace Foam
{
int GAMGAgglomeration::debug = 0;
}
using namespace Foam;

namespace
{
class FineMesh : public lduMesh
{
    lduAddressing a_;

public:
    FineMesh(const labelList &l, const labelList &u, label n) : a_(l, u, n) {}
    virtual const lduAddressing &lduAddr() const { return a_; }
};
} // namespace

extern "C" {
/* nSteps (1 or 2) consecutive pairing steps with given restrict maps (map0: fine -> level 1, map1: level 1 ->
 * level 2).  With nSteps == 2 the second level is folded into the first (combineLevels(1)).  Outputs describe
 * level 0 afterwards: composed restrict map [nCells], face restrict map and flip [nFaces], coarse
 * owner/neighbour [*nCoarseFaces].  Returns the number of coarse cells, or -1 on a FatalError. */
int ref_coarse_levels(int nSteps, int nCells, int nFaces, const int *lower, const int *upper, const int *map0,
                      int nCoarse0, const int *map1, int nCoarse1, int *restrictOut, int *faceRestrictOut,
                      unsigned char *flipOut, int *nCoarseFaces, int *coarseOwner, int *coarseNeighbour)
{
    try {
        labelList l(lower, nFaces), u(upper, nFaces);
        FineMesh fine(l, u, nCells);
        GAMGAgglomeration agg(fine, 4);
        agg.useAtomic_ = false; // the sorted (non-atomic) addressing path, as pairGAMGAgglomerate.C:66-77 builds it
        agg.nCells_[0] = nCoarse0;
        agg.restrictAddressingHost_.set(0, new labelField(map0, nCells));
        agg.buildFullRestrictAddr(labelgpuList(agg.restrictAddressingHost_[0]), 0);
        agg.agglomerateLduAddressing(0);
        if (nSteps == 2) {
            agg.nCells_[1] = nCoarse1;
            agg.restrictAddressingHost_.set(1, new labelField(map1, nCoarse0));
            agg.buildFullRestrictAddr(labelgpuList(agg.restrictAddressingHost_[1]), 1);
            agg.agglomerateLduAddressing(1);
            agg.combineLevels(1);
        }
        const labelField &r = agg.restrictAddressingHost_[0];
        for (label i = 0; i < nCells; i++) restrictOut[i] = r[i];
        const labelList &fr = agg.faceRestrictAddressingHost_[0];
        const boolList &ff = agg.faceFlipMapHost_[0];
        for (label f = 0; f < nFaces; f++) {
            faceRestrictOut[f] = fr[f];
            flipOut[f] = ff[f];
        }
        const lduAddressing &ca = agg.meshLevels_[0].lduAddr();
        *nCoarseFaces = ca.upperAddrHost().size();
        for (label f = 0; f < *nCoarseFaces; f++) {
            coarseOwner[f] = ca.lowerAddrHost()[f];
            coarseNeighbour[f] = ca.upperAddrHost()[f];
        }
        return agg.nCells_[0];
    } catch (const std::exception &) {
        return -1;
    }
}
}
