/* Internal structs of the CPU oracle (test infrastructure only; see ldu_oracle.h). */
#ifndef LDU_ORACLE_INTERNAL_H
#define LDU_ORACLE_INTERNAL_H
#include "ldu_oracle.h"

struct orc_addr {
    int nCells, nFaces;
    int *l, *u;        /* lowerAddr (owner), upperAddr (neighbour) */
    int *ownerStart;   /* [nCells+1] lduAddressing.C:202-267 */
    int *losort;       /* [nFaces]   lduAddressing.C:169-199 */
    int *losortStart;  /* [nCells+1] lduAddressing.C:270-344 */
    int nPatches;      /* coupled (processor / cyclic) patches only */
    int *patchStart;   /* [nPatches+1] offsets into the flat patch arrays */
    int *faceCells;    /* flat */
    int *neighbRank;   /* [nPatches] or NULL */
};

struct orc_matrix {
    const orc_addr *a;
    const double *diag, *upper, *lower; /* lower aliases upper when symmetric */
    int symmetric;
    const double *bou, *intc; /* interfaceBouCoeffs / interfaceIntCoeffs, flat */
};

double orc_gsum(const double *x, int n, const orc_comm *comm);
double orc_gsumprod(const double *x, const double *y, int n, const orc_comm *comm);
double orc_gsummag(const double *x, int n, const orc_comm *comm);
int orc_precond_kind(const char *name, char *printed);
double *orc_halo_exchange(const orc_addr *a, const double *psi, const orc_comm *comm);

#endif
