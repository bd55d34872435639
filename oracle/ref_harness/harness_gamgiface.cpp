/*
 * harness_gamgiface.cpp -- runs the REFERENCE'S OWN multi-rank coarse-level construction on the CPU: the
 * ranks of a decomposed case live in one process and exchange their restrict maps through a mailbox behind
 * processorLduInterface::send / receive.  TEST INFRASTRUCTURE ONLY.  Included by path from /root/reference:
 *   GAMG/GAMGAgglomerations/GAMGAgglomeration/GAMGAgglomerateLduAddressing.C
 *       agglomerateLduAddressing :245-603 with its interface branch (:466-556), combineLevels :606-765
 *   GAMG/interfaces/GAMGInterface/GAMGInterface.C, GAMGInterfaceNew.C, GAMGInterfaceTemplates.C, GAMGInterfaceF.H
 *       updateAddressing, combine, interfaceInternalField, agglomerateCoeffs (+ its functor), New
 *   GAMG/interfaces/processorGAMGInterface/processorGAMGInterface.C
 *       coarse patch faces = unique (master coarse cell, slave coarse cell) pairs in order of appearance
 *       :52-139; initInternalFieldTransfer / internalFieldTransfer :196-226
 * against oracle/ref_harness/shim_gamgiface/ (+ shim_gamgaddr/ for the containers and GAMGAgglomeration).
 */
#define SHIM_REAL_GAMG_INTERFACE
#include "GAMGAgglomeration.H" /* shim */
#include "GAMGInterface.H"     /* shim: class declaration */
#include "processorGAMGInterface.H"

#include "GAMGAgglomerateLduAddressing.C" /* reference */
#include "GAMGInterface.C"                /* reference */
#include "GAMGInterfaceNew.C"             /* reference */
#include "GAMGInterfaceTemplates.C"       /* reference */
#include "processorGAMGInterface.C"       /* reference */

#include <memory>

namespace Foam
{
int GAMGAgglomeration::debug = 0;
label UPstream::warnComm = -1;
} // namespace Foam
using namespace Foam;

namespace
{
// the finest-level coupled patch (processorFvPatch in the reference): face cells and the rank pair
class FineProc : public lduInterface, public processorLduInterface
{
    labelList cellsHost_;
    labelgpuList cells_;
    int me_, nbr_;
    tensorField T_;
    static const word typeName_;

public:
    FineProc(const int *fc, label n, int me, int nbr) : cellsHost_(fc, n), cells_(cellsHost_), me_(me), nbr_(nbr) {}
    virtual const word &type() const { return typeName_; }
    virtual const labelgpuList &faceCells() const { return cells_; }
    virtual tmp<labelField> interfaceInternalField(const labelUList &iF) const // lduInterface.C:45-60
    {
        tmp<labelField> t(new labelField(cellsHost_.size()));
        forAll(cellsHost_, i) t()[i] = iF[cellsHost_[i]];
        return t;
    }
    virtual void initInternalFieldTransfer(Pstream::commsTypes ct, const labelUList &iF) const
    {
        send(ct, interfaceInternalField(iF)()); // processorFvPatch.C:128-135
    }
    virtual tmp<labelField> internalFieldTransfer(Pstream::commsTypes ct, const labelUList &) const
    {
        tmp<labelField> t(new labelField(cellsHost_.size()));
        receive<label>(ct, t());
        return t;
    }
    virtual int comm() const { return 0; }
    virtual int myProcNo() const { return me_; }
    virtual int neighbProcNo() const { return nbr_; }
    virtual const tensorField &forwardT() const { return T_; }
    virtual int tag() const { return 1; }
};
const word FineProc::typeName_("processor");

class FineMesh : public lduMesh
{
    lduAddressing a_;

public:
    FineMesh(const labelList &l, const labelList &u, label n) : a_(l, u, n) {}
    virtual const lduAddressing &lduAddr() const { return a_; }
};

struct Rank {
    std::unique_ptr<FineMesh> mesh;
    std::unique_ptr<GAMGAgglomeration> agg;
    std::vector<std::unique_ptr<FineProc>> patches;
};
struct Case {
    std::vector<Rank> ranks;
};
std::vector<std::unique_ptr<Case>> g_cases;
const int MAX_LEVELS = 8;
} // namespace

extern "C" {
int ref_ia_create(int nRanks)
{
    g_cases.emplace_back(new Case);
    g_cases.back()->ranks.resize((size_t)nRanks);
    return (int)g_cases.size() - 1;
}

void ref_ia_destroy(int h) { g_cases[(size_t)h].reset(); }

int ref_ia_set_rank(int h, int r, int nCells, int nFaces, const int *lower, const int *upper, int nPatches,
                    const int *patchStart, const int *faceCells, const int *neighbRank)
{
    Rank &R = g_cases[(size_t)h]->ranks[(size_t)r];
    labelList l(lower, nFaces), u(upper, nFaces);
    R.mesh.reset(new FineMesh(l, u, nCells));
    R.agg.reset(new GAMGAgglomeration(*R.mesh, MAX_LEVELS));
    R.agg->useAtomic_ = false;
    R.agg->meshInterfaces_ = lduInterfacePtrsList(nPatches);
    for (int p = 0; p < nPatches; p++) {
        R.patches.emplace_back(
            new FineProc(faceCells + patchStart[p], patchStart[p + 1] - patchStart[p], r, neighbRank[p]));
        R.agg->meshInterfaces_.set(p, R.patches.back().get());
    }
    return 0;
}

/* restrict map of rank r from level `level` to level+1 */
int ref_ia_set_map(int h, int level, int r, const int *map, int nFine, int nCoarse)
{
    GAMGAgglomeration &a = *g_cases[(size_t)h]->ranks[(size_t)r].agg;
    a.nCells_[level] = nCoarse;
    a.restrictAddressingHost_.set(level, new labelField(map, nFine));
    a.buildFullRestrictAddr(labelgpuList(a.restrictAddressingHost_[level]), level);
    return 0;
}

/* every rank posts its restrict map on its interfaces (the sends the reference issues at the top of the
 * interface branch, done for all ranks first because the ranks run one after the other here), then each
 * rank runs agglomerateLduAddressing(level) */
int ref_ia_agglomerate(int h, int level)
{
    try {
        Case &C = *g_cases[(size_t)h];
        for (Rank &R : C.ranks) {
            const lduInterfacePtrsList &ifs = R.agg->interfaceLevel(level);
            forAll(ifs, i) if (ifs.set(i))
                ifs[i].initInternalFieldTransfer(Pstream::nonBlocking, R.agg->restrictAddressingHost(level));
        }
        for (Rank &R : C.ranks) R.agg->agglomerateLduAddressing(level);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

int ref_ia_combine(int h, int level)
{
    try {
        for (Rank &R : g_cases[(size_t)h]->ranks) R.agg->combineLevels(level);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

/* sizes of what level `level` of rank r produced: [nCoarseCells, nCoarseFaces, nPatches, coarse patch sizes...] */
int ref_ia_sizes(int h, int level, int r, int *out)
{
    GAMGAgglomeration &a = *g_cases[(size_t)h]->ranks[(size_t)r].agg;
    out[0] = a.nCells_[level];
    out[1] = a.meshLevels_[level].lduAddr().upperAddrHost().size();
    const labelList &np = a.nPatchFaces_[level];
    out[2] = np.size();
    forAll(np, p) out[3 + p] = np[p];
    return 0;
}

/* coarse owner / neighbour, fine-face restrict map + flip, per patch: coarse face cells (concatenated) and the
 * fine patch face -> coarse patch face map (concatenated, patch-local numbers) */
int ref_ia_get(int h, int level, int r, int *coarseOwner, int *coarseNeighbour, int *faceRestrict, unsigned char *flip,
               int *coarseFaceCells, int *patchFaceRestrict)
{
    GAMGAgglomeration &a = *g_cases[(size_t)h]->ranks[(size_t)r].agg;
    const lduAddressing &ca = a.meshLevels_[level].lduAddr();
    forAll(ca.upperAddrHost(), f)
    {
        coarseOwner[f] = ca.lowerAddrHost()[f];
        coarseNeighbour[f] = ca.upperAddrHost()[f];
    }
    const labelList &fr = a.faceRestrictAddressingHost_[level];
    const boolList &ff = a.faceFlipMapHost_[level];
    forAll(fr, f)
    {
        faceRestrict[f] = fr[f];
        flip[f] = ff[f];
    }
    const lduInterfacePtrsList &ci = a.meshLevels_[level].rawInterfaces();
    const labelListList &pfr = a.patchFaceRestrictAddressingHost_[level];
    label nc = 0, nf = 0;
    forAll(ci, p)
    {
        if (!ci.set(p)) continue;
        const GAMGInterface &gi = refCast<const GAMGInterface>(ci[p]);
        forAll(gi.faceCellsHost(), i) coarseFaceCells[nc++] = gi.faceCellsHost()[i];
        forAll(pfr[p], i) patchFaceRestrict[nf++] = pfr[p][i];
    }
    return 0;
}

/* GAMGInterface::agglomerateCoeffs of coarse patch p created by level `level` (GAMGInterface.C:120-172) */
int ref_ia_coeffs(int h, int level, int r, int p, const double *fine, int nFine, double *coarse)
{
    try {
        GAMGAgglomeration &a = *g_cases[(size_t)h]->ranks[(size_t)r].agg;
        const GAMGInterface &gi = refCast<const GAMGInterface>(a.meshLevels_[level].rawInterfaces()[p]);
        scalargpuField f(fine, nFine);
        tmp<scalargpuField> t = gi.agglomerateCoeffs(f);
        for (label i = 0; i < t().size(); i++) coarse[i] = t()[i];
        return t().size();
    } catch (const std::exception &) {
        return -1;
    }
}
}
