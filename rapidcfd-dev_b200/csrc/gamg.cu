// gamg.cu -- GAMG: cached pair agglomeration + device V-cycle.
// (entry points are defined here; the device V-cycle is filled in below)
#include "ldu.h"
#include "solvers.h"

struct b200ldu_gamg {
    b200ldu_addr *finest = nullptr;
    int nLevels = 0;
};

int gamg_solve(Solve &S, b200ldu_gamg *g, const char *smoother)
{
    b200_set_error("GAMG: device V-cycle not built yet");
    return B200LDU_ENOSOLVER;
}

extern "C" int b200ldu_gamg_create(b200ldu_addr *a, const double *faceWeights_h, int nCellsInCoarsestLevel,
                                   int mergeLevels, int *forward, b200ldu_gamg **out)
{
    b200_set_error("GAMG: not built yet");
    return B200LDU_ENOSOLVER;
}
extern "C" int b200ldu_gamg_destroy(b200ldu_gamg *g)
{
    delete g;
    return B200LDU_OK;
}
extern "C" int b200ldu_gamg_nlevels(const b200ldu_gamg *g) { return g ? g->nLevels : 0; }
extern "C" int b200ldu_gamg_level_size(const b200ldu_gamg *g, int lev, int *nCells, int *nFaces)
{
    return B200LDU_EINVAL;
}
extern "C" int b200ldu_gamg_restrict_addr(const b200ldu_gamg *g, int lev, int *out_h) { return B200LDU_EINVAL; }
