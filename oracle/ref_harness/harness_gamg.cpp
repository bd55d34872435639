/*
 * harness_gamg.cpp -- runs the REFERENCE'S OWN pair agglomeration on the CPU.  TEST INFRASTRUCTURE ONLY.
 * Includes, by path from /root/reference (through the symlink oracle/_ref/inc_gamg/):
 *   LDU/solvers/GAMG/GAMGAgglomerations/pairGAMGAgglomeration/pairGAMGAgglomerate.C
 * against oracle/ref_harness/shim_gamg/.  See oracle/Makefile target `ref`.
 */
#include "pairGAMGAgglomeration.H" /* the shim */

#include "pairGAMGAgglomerate.C" /* reference */

namespace Foam
{
bool pairGAMGAgglomeration::forward_(true); // pairGAMGAgglomeration.C:33
// members used only by the level loop of the reference file (never called here)
bool pairGAMGAgglomeration::useAtomic() const { abort(); }
void pairGAMGAgglomeration::buildFullRestrictAddr(const labelgpuList &, label) { abort(); }
const lduMesh &pairGAMGAgglomeration::meshLevel(label) const { abort(); }
void pairGAMGAgglomeration::agglomerateLduAddressing(label) { abort(); }
void pairGAMGAgglomeration::combineLevels(label) { abort(); }
void pairGAMGAgglomeration::compactLevels(label) { abort(); }
bool pairGAMGAgglomeration::continueAgglomerating(label) const { abort(); }
} // namespace Foam

extern "C" {
/* one pairing step: map[nCells] fine -> coarse, returns nCoarseCells; *forward is the static
 * direction flag before the call and is updated like the reference updates it */
int ref_pair_agglomerate(int nCells, int nFaces, const int *lower, const int *upper, const double *faceWeights,
                         int *forward, int *map)
{
    using namespace Foam;
    lduAddressing addr(lower, upper, nFaces, nCells);
    scalarField w(faceWeights, nFaces);
    pairGAMGAgglomeration::forward_ = (*forward != 0);
    label nCoarse = -1;
    tmp<labelField> t = pairGAMGAgglomeration::agglomerate(nCoarse, addr, w);
    for (label i = 0; i < nCells; i++) map[i] = t()[i];
    *forward = pairGAMGAgglomeration::forward_ ? 1 : 0;
    return nCoarse;
}
}
