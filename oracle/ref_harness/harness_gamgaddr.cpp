/*
 * harness_gamgaddr.cpp -- runs the REFERENCE'S OWN coarse-addressing construction and level
 * combination on the CPU.  TEST INFRASTRUCTURE ONLY.  Included by path from /root/reference:
 *   GAMG/GAMGAgglomerations/GAMGAgglomeration/GAMGAgglomerateLduAddressing.C
 *       agglomerateLduAddressing :245-603, combineLevels :606-765 (+ createSort/createTarget/buildFull*)
 *   GAMG/GAMGAgglomerations/GAMGAgglomeration/GAMGAgglomerationF.H, GAMG/GAMGSolverAgglomerateMatrixF.H
 *       the restriction / coarse-coefficient functors, launched over the reference-built sorted addressing
 *       the way GAMGAgglomerationTemplates.C:35-61 and GAMGSolverAgglomerateMatrix.C:183-320 launch them
 * against oracle/ref_harness/shim_gamgaddr/.
 */
#include "GAMGAgglomeration.H" /* shim */

#include "GAMGAgglomerateLduAddressing.C" /* reference */
#include "GAMGAgglomerationF.H"           /* reference: GAMG::restrict (segmented sum over the sorted addressing) */
#include "GAMGSolverAgglomerateMatrixF.H" /* reference: GAMG::symAgglomerate, asymAgglomerate, diag*Agglomerate */
#include "GAMGAgglomerateF.H"             /* reference: GAMG::negative, nonNegative, faceToDiag */

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
                      unsigned char *flipOut, int *nCoarseFaces, int *coarseOwner, int *coarseNeighbour,
                      const double *fineDiag, const double *fineUpper, const double *fineLower, double *coarseDiag,
                      double *coarseUpper, double *coarseLower)
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
        if (fineDiag) { // coarse coefficients of level 0 (single step only)
            const labelgpuList &cs = agg.restrictSortAddressing_[0], &ct = agg.restrictTargetAddressing_[0],
                               &cts = agg.restrictTargetStartAddressing_[0];
            auto seg = [](const labelgpuList &ts) {
                return thrust::make_zip_iterator(thrust::make_tuple(ts.begin(), ts.begin() + 1));
            };
            // restrictField: GAMGAgglomerationTemplates.C:35-61
            for (label i = 0; i < agg.nCells_[0]; i++) coarseDiag[i] = 0;
            thrust::transform(cts.begin(), cts.end() - 1, cts.begin() + 1,
                              thrust::make_permutation_iterator(coarseDiag, ct.begin()),
                              GAMG::restrict<scalar>(fineDiag, cs.data()));
            const labelgpuList &fs = agg.faceRestrictSortAddressing_[0], &ft = agg.faceRestrictTargetAddressing_[0],
                               &fts = agg.faceRestrictTargetStartAddressing_[0];
            const label nT = ft.size(), nCF = agg.nFaces_[0];
            for (label i = 0; i < nCF; i++) coarseUpper[i] = 0;
            auto diagIter = thrust::make_permutation_iterator(
                coarseDiag, thrust::make_transform_iterator(ft.begin(), GAMG::faceToDiag()));
            if (fineLower) { // GAMGSolverAgglomerateMatrix.C:196-268
                for (label i = 0; i < nCF; i++) coarseLower[i] = 0;
                std::vector<char> flips(agg.faceFlipMapHost_[0].begin(), agg.faceFlipMapHost_[0].end());
                static_assert(sizeof(bool) == 1, "flip map is read as bool");
                auto lu = thrust::make_zip_iterator(thrust::make_tuple(
                    thrust::make_permutation_iterator(coarseUpper, ft.begin()),
                    thrust::make_permutation_iterator(coarseLower, ft.begin())));
                thrust::transform_if(lu, lu + nT, seg(fts), ft.begin(), lu,
                                     GAMG::asymAgglomerate(fineUpper, fineLower,
                                                           reinterpret_cast<const bool *>(flips.data()), fs.data()),
                                     GAMG::nonNegative());
                thrust::transform_if(diagIter, diagIter + nT, seg(fts), ft.begin(), diagIter,
                                     GAMG::diagAsymAgglomerate(fineUpper, fineLower, fs.data()), GAMG::negative());
            } else { // :270-318
                auto up = thrust::make_permutation_iterator(coarseUpper, ft.begin());
                thrust::transform_if(up, up + nT, seg(fts), ft.begin(), up, GAMG::symAgglomerate(fineUpper, fs.data()),
                                     GAMG::nonNegative());
                thrust::transform_if(diagIter, diagIter + nT, seg(fts), ft.begin(), diagIter,
                                     GAMG::diagSymAgglomerate(fineUpper, fs.data()), GAMG::negative());
            }
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
