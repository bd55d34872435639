/*
 * harness_limiters.cpp -- the REFERENCE'S OWN limiter functions evaluated on the host: TEST INFRASTRUCTURE.
 * Included by path from /root/reference (symlinks in oracle/_ref/inc_limiters/ while compiling):
 *   FV/interpolation/surfaceInterpolation/limitedSchemes/LimitedScheme/NVDTVD.H      r() :99-127
 *   .../limitedLinear/limitedLinear.H   limitedLinearLimiter<NVDTVD>::limiter :83-101, twoByk_ :74-77
 *   .../vanLeer/vanLeer.H               vanLeerLimiter<NVDTVD>::limiter :66-85
 *   .../Minmod/Minmod.H                 MinmodLimiter<NVDTVD>::limiter :66-85
 * against shim_limiters/vector.H.  The loop over the faces restates LimitedSchemeCalcLimiterFunctor
 * (LimitedScheme.H: limiter(cdWeight, faceFlux, phiP, phiN, gradcP, gradcN, C_N - C_P)); the weights formula
 * restates limitedSurfaceInterpolationSchemeWeightsFunctor (limitedSurfaceInterpolationScheme.C:155-163), which
 * cannot be included without the whole class.
 */
#include "vector.H"

#include "NVDTVD.H"
#include "limitedLinear.H"
#include "vanLeer.H"
#include "Minmod.H"

using namespace Foam;

template <class Lim>
static void run(Lim &lim, int nF, const int *l, const int *u, const double *cd, const double *flux, const double *vf,
                const double *g, const double *C, double *out)
{
    for (int f = 0; f < nF; f++) {
        const int o = l[f], n = u[f];
        const vector gP{g[3 * o], g[3 * o + 1], g[3 * o + 2]}, gN{g[3 * n], g[3 * n + 1], g[3 * n + 2]};
        const vector d{C[3 * n] - C[3 * o], C[3 * n + 1] - C[3 * o + 1], C[3 * n + 2] - C[3 * o + 2]};
        out[f] = lim.limiter(cd[f], flux[f], vf[o], vf[n], gP, gN, d);
    }
}

extern "C" int ref_limiter(int scheme, double k, int nF, const int *l, const int *u, const double *cd, const double *flux,
                           const double *vf, const double *gradc, const double *C, double *out)
{
    Istream is{k};
    try {
        if (scheme == 2) {
            limitedLinearLimiter<NVDTVD> lim(is);
            run(lim, nF, l, u, cd, flux, vf, gradc, C, out);
        } else if (scheme == 3) {
            vanLeerLimiter<NVDTVD> lim(is);
            run(lim, nF, l, u, cd, flux, vf, gradc, C, out);
        } else if (scheme == 4) {
            MinmodLimiter<NVDTVD> lim(is);
            run(lim, nF, l, u, cd, flux, vf, gradc, C, out);
        } else
            return -1;
    } catch (const std::runtime_error &) {
        return -2;
    }
    return 0;
}
