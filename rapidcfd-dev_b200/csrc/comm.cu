// comm.cu -- NCCL plumbing: one communicator per context (one process per GPU).
// Replaces the reference's Pstream-over-MPI layer for the hot path: processor-patch
// halo exchange (src/Pstream/mpi/UOPwrite.C:73-122, UIPread.C:260-316,
// LDU/lduAddressing/lduInterface/processorLduInterfaceTemplates.C:128-298) and the
// scalar all-reduces behind gSumProd/gSumMag/gAverage
// (src/Pstream/mpi/allReduceTemplates.C:197).  Everything is enqueued on the context's
// stream; nothing is staged through host memory.
#include <nccl.h>

#include "comm.h"

#define NCCL_TRY(expr)                                                                      \
    do {                                                                                    \
        ncclResult_t _r = (expr);                                                           \
        if (_r != ncclSuccess) {                                                            \
            b200_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, ncclGetErrorString(_r)); \
            return B200LDU_ENCCL;                                                           \
        }                                                                                   \
    } while (0)

extern "C" int b200ldu_comm_unique_id(void *out128)
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_TRY(ncclGetUniqueId(&id));
    memcpy(out128, &id, 128);
    return B200LDU_OK;
}

extern "C" int b200ldu_comm_init(b200ldu_ctx *ctx, const void *id128, int rank, int nRanks)
{
    if (!ctx || nRanks < 1 || rank < 0 || rank >= nRanks) {
        b200_set_error("comm_init: bad arguments");
        return B200LDU_EINVAL;
    }
    ctx->rank = rank;
    ctx->nRanks = nRanks;
    if (nRanks == 1) return B200LDU_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    NCCL_TRY(ncclCommInitRank(&comm, nRanks, id, rank));
    ctx->nccl = (void *)comm;
    return B200LDU_OK;
}

int comm_destroy(b200ldu_ctx *ctx)
{
    if (ctx->nccl) {
        ncclCommDestroy((ncclComm_t)ctx->nccl);
        ctx->nccl = nullptr;
    }
    return B200LDU_OK;
}

int comm_allreduce_sum(b200ldu_ctx *ctx, double *d_buf, int n)
{
    if (ctx->nRanks == 1) return B200LDU_OK;
    NCCL_TRY(ncclAllReduce(d_buf, d_buf, n, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
    return B200LDU_OK;
}

__global__ void pack_kernel(int n, const int *__restrict__ rows, const double *__restrict__ x,
                            double *__restrict__ send, const int *stop)
{
    if (stop && *stop) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) send[i] = x[rows[i]];
}

// x is a banded vector of vecLen doubles; received neighbour values land in its tail
// [nPad, nPad + nRecv) where the band halo lists point.
int comm_halo_exchange(b200ldu_addr *a, double *x, double *sendBuf, const int *stop)
{
    const int nRecv = a->L.nRecv;
    if (nRecv == 0) return B200LDU_OK;
    b200ldu_ctx *ctx = a->ctx;
    pack_kernel<<<(nRecv + 255) / 256, 256, 0, ctx->stream>>>(nRecv, a->d_sendRows, x, sendBuf, stop);
    ctx->launches++;
    KERNEL_CHECK();
    bool remote = false;
    for (int p = 0; p < a->nPatches; p++)
        if (a->neighbRank[p] != ctx->rank) remote = true;
    if (remote && !ctx->nccl) {
        b200_set_error("halo exchange: processor patches present but no communicator (b200ldu_comm_init)");
        return B200LDU_ENCCL;
    }
    if (remote) NCCL_TRY(ncclGroupStart());
    for (int p = 0; p < a->nPatches; p++) {
        int s = a->patchStart[p], n = a->patchStart[p + 1] - s;
        int nb = a->neighbRank[p];
        if (nb == ctx->rank) {
            b200_set_error("halo exchange: cyclic (same-rank) interfaces are not supported yet");
            return B200LDU_EINVAL;
        }
        NCCL_TRY(ncclSend(sendBuf + s, n, ncclDouble, nb, (ncclComm_t)ctx->nccl, ctx->stream));
        NCCL_TRY(ncclRecv(x + a->L.nPad + s, n, ncclDouble, nb, (ncclComm_t)ctx->nccl, ctx->stream));
    }
    if (remote) NCCL_TRY(ncclGroupEnd());
    return B200LDU_OK;
}
