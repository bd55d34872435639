/*
 * harness_procfield.cpp -- runs the REFERENCE'S OWN finest-level coupled-interface update on the CPU, all ranks of a
 * decomposed case in one process: lduMatrix::initMatrixInterfaces posts every patch's send (and registers its receive),
 * lduMatrix::updateMatrixInterfaces completes them and applies result[faceCells] -= coeffs*psi_neighbour.  TEST
 * INFRASTRUCTURE ONLY.  Included by path from /root/reference:
 *   LDU/lduMatrix/lduMatrixUpdateMatrixInterfaces.C                       :30-276
 *   FV/fields/fvPatchFields/constraint/processor/processorFvPatchScalarField.C  :37-172
 *   LDU/lduAddressing/lduAddressingFunctors.H                             matrixPatchOperation, matrixInterfaceFunctor
 *   GAMG/interfaceFields/processorGAMGInterfaceField/processorGAMGInterfaceField.H/.C  :94-248 (coarse levels)
 *   GAMG/interfaces/GAMGInterface/GAMGInterfaceFunctors.H                 GAMGUpdateInterfaceMatrix
 *   FV/fields/fvPatchFields/constraint/cyclic/cyclicFvPatchField.C        updateInterfaceMatrix :212-231 (cyclic pairs)
 * against oracle/ref_harness/shim_procfield/ (+ shim/foam_shim.h).
 */
#include "procfield_shim.h"

#include <algorithm>
#include <memory>

#include "lduMatrixUpdateMatrixInterfaces.C" /* reference */
#include "processorFvPatchScalarField.C"     /* reference */
#include "cyclicFvPatchField.C"              /* reference: updateInterfaceMatrix :212-231 */
#include "processorGAMGInterfaceField.H"     /* reference: class declaration */
#include "processorGAMGInterfaceField.C"     /* reference */

namespace Foam
{
UPstream::commsTypes UPstream::defaultCommsType = UPstream::nonBlocking;
bool UPstream::floatTransfer = false, UPstream::gpuDirectTransfer = false;
label UPstream::nPollProcInterfaces = 0;
label UPstream::warnComm = -1;
const char *UPstream::commsTypeNames[3] = {"blocking", "scheduled", "nonBlocking"};
template <> int processorFvPatchField<scalar>::debug = 0;
template <> const char *cyclicFvPatchField<scalar>::typeName = "cyclic";
int lduMatrix::debug = 0;
int lduMatrixSolutionCache::favourSpeed = 0;
} // namespace Foam
using namespace Foam;

namespace
{
struct Rank {
    fvMeshStub mesh;
    std::vector<std::unique_ptr<processorFvPatch>> patches;
    std::vector<std::unique_ptr<processorFvPatchField<scalar>>> fields;
    std::vector<std::unique_ptr<cyclicFvPatch>> cycPatches;               // neighbRank < 0: cyclic partner -(q+1)
    std::vector<std::unique_ptr<cyclicFvPatchField<scalar>>> cycFields;
    std::vector<const lduInterfaceField *> ifaceOf;
    std::vector<std::unique_ptr<processorGAMGInterface>> gamgPatches;     // coarse-level variant
    std::vector<std::unique_ptr<processorGAMGInterfaceField>> gamgFields;
    std::vector<scalargpuField> coeffs;
    lduMatrix M;
    lduInterfaceFieldPtrsList interfaces;
    FieldField<gpuField, scalar> coupleCoeffs;
    scalargpuField psi, result;
};
std::vector<std::unique_ptr<Rank>> g_ranks;
} // namespace

extern "C" {
void ref_pf_reset(int nRanks)
{
    g_ranks.clear();
    for (int r = 0; r < nRanks; r++) g_ranks.emplace_back(new Rank);
    Mail::box().clear();
    Mail::pending().clear();
}

/* one rank: cells, its coupled patches (face cells, neighbour ranks, coefficients), psi and the vector to update */
void ref_pf_set_rank(int r, int nCells, int nPatches, const int *patchStart, const int *faceCells, const int *neighbRank,
                     const double *coeffs, const double *psi, const double *result, int gamgLevel)
{
    Rank &R = *g_ranks[(size_t)r];
    lduAddressing &a = R.mesh.addr_;
    a.nCells_ = nCells;
    R.interfaces = lduInterfaceFieldPtrsList(nPatches);
    R.coupleCoeffs = FieldField<gpuField, scalar>(nPatches);
    R.coeffs.resize((size_t)nPatches);
    for (int p = 0; p < nPatches; p++) {
        const int s = patchStart[p], np = patchStart[p + 1] - s;
        std::vector<label> order(np), cells, start;
        for (int i = 0; i < np; i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](label x, label y) { return faceCells[s + x] < faceCells[s + y]; });
        for (int k = 0; k < np; k++) {
            const label c = faceCells[s + order[k]];
            if (cells.empty() || cells.back() != c) {
                cells.push_back(c);
                start.push_back(k);
            }
        }
        start.push_back(np);
        a.patchCellsV_.push_back(labelgpuList(cells.data(), (label)cells.size()));
        a.patchSortV_.push_back(labelgpuList(order.data(), np));
        a.patchSortStartV_.push_back(labelgpuList(start.data(), (label)start.size()));
        R.coeffs[(size_t)p] = scalargpuField(coeffs + s, np);
        if (gamgLevel) { // processorGAMGInterface with the cell-sorted face lists of GAMGInterface::updateAddressing
            processorGAMGInterface *gp = new processorGAMGInterface;
            gp->faceCells_ = labelgpuList(faceCells + s, np);
            gp->sortCells_ = labelgpuList(cells.data(), (label)cells.size());
            gp->cellFaces_ = labelgpuList(order.data(), np);
            gp->cellFacesStart_ = labelgpuList(start.data(), (label)start.size());
            gp->myProcNo_ = r;
            gp->neighbProcNo_ = neighbRank[p];
            gp->tag_ = 1;
            R.gamgPatches.emplace_back(gp);
            R.gamgFields.emplace_back(new processorGAMGInterfaceField(*gp, false, 0));
            continue;
        }
        if (neighbRank[p] < 0) { // cyclic: the neighbour values are psi at the partner patch's face cells
            cyclicFvPatch *cp = new cyclicFvPatch;
            cp->mesh_ = &R.mesh;
            cp->index_ = p;
            cp->faceCells_ = labelgpuList(faceCells + s, np);
            cp->nbrID_ = -neighbRank[p] - 1;
            R.cycPatches.emplace_back(cp);
            R.cycFields.emplace_back(new cyclicFvPatchField<scalar>(*cp, 0));
            R.ifaceOf.push_back(R.cycFields.back().get());
            continue;
        }
        processorFvPatch *pp = new processorFvPatch;
        pp->mesh_ = &R.mesh;
        pp->index_ = p;
        pp->faceCells_ = labelgpuList(faceCells + s, np);
        pp->myProcNo_ = r;
        pp->neighbProcNo_ = neighbRank[p];
        pp->tag_ = 1;
        R.patches.emplace_back(pp);
        R.fields.emplace_back(new processorFvPatchField<scalar>(*pp));
        R.ifaceOf.push_back(R.fields.back().get());
    }
    for (auto &cp : R.cycPatches) // partner patches: located by their index among this rank's patches
        for (auto &cq : R.cycPatches)
            if (cq->index_ == cp->nbrID_) cp->nbr_ = cq.get();
    for (int p = 0; p < nPatches; p++) { // after the vectors stopped growing
        R.interfaces.set(p, gamgLevel ? static_cast<const lduInterfaceField *>(R.gamgFields[(size_t)p].get()) : R.ifaceOf[(size_t)p]);
        R.coupleCoeffs.setPtr(p, &R.coeffs[(size_t)p]);
    }
    R.M.addr_ = &a;
    R.psi = scalargpuField(psi, nCells);
    R.result = scalargpuField(result, nCells);
}

/* commsType: 0 blocking, 2 nonBlocking (Pstream::defaultCommsType); negate as the smoothers pass it */
int ref_pf_update(int commsType, int negate, int nPoll, int gpuDirect)
{
    try {
        UPstream::defaultCommsType = (UPstream::commsTypes)commsType;
        UPstream::nPollProcInterfaces = nPoll;
        UPstream::gpuDirectTransfer = gpuDirect != 0;
        for (size_t r = 0; r < g_ranks.size(); r++) {
            Mail::me() = (int)r;
            Rank &R = *g_ranks[r];
            R.M.initMatrixInterfaces(R.coupleCoeffs, R.interfaces, R.psi, R.result, 0, negate != 0);
        }
        for (size_t r = 0; r < g_ranks.size(); r++) {
            Mail::me() = (int)r;
            Rank &R = *g_ranks[r];
            R.M.updateMatrixInterfaces(R.coupleCoeffs, R.interfaces, R.psi, R.result, 0, negate != 0);
        }
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

void ref_pf_get(int r, double *result)
{
    const Rank &R = *g_ranks[(size_t)r];
    std::copy(R.result.begin(), R.result.end(), result);
}
}
