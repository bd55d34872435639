// comm.h -- internal interface of comm.cu
#pragma once
#include "internal.h"
int comm_destroy(b200ldu_ctx *ctx);
int comm_allreduce_sum(b200ldu_ctx *ctx, double *d_buf, int n);
int comm_halo_exchange(b200ldu_addr *a, double *x, double *sendBuf, const int *stop, int *usedP2P);
int comm_addr_setup(b200ldu_addr *a);
P2PRed comm_p2p_red(b200ldu_ctx *ctx);
int comm_exchange_patch_ints(b200ldu_ctx *ctx, int nPatches, const int *patchStart, const int *neighbRank,
                             const int *send, int *recv);
int comm_allgather_host(b200ldu_ctx *ctx, const double *mine, int n, double *all);
int comm_allgather_dev(b200ldu_ctx *ctx, const double *d_mine, int n, double *d_all);
int comm_exchange_patch_field(b200ldu_addr *a, int nComp, const double *send_d, double *recv_d);
