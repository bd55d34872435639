"""CPU restatement (numpy) of the reference's limited interpolation schemes on the internal faces -- TEST
INFRASTRUCTURE (the checker of b200ldu_fv_limiter / b200ldu_fv_limited_weights; never imported by the product).

  NVDTVD::r                       FV/interpolation/surfaceInterpolation/limitedSchemes/LimitedScheme/NVDTVD.H:99-127
  LimitedScheme::calcLimiter      .../LimitedScheme/LimitedScheme.C:60-140 (d = C[nei] - C[own], gradc = fvc::grad(vf))
  limitedLinear / vanLeer / Minmod limiter functions   limitedLinear.H:74-101, vanLeer.H:66-85, Minmod.H:66-85
  upwind                          upwind.H:103-123 (limiter 0, weights pos(faceFlux))
  weights                         limitedSurfaceInterpolationScheme.C:155-212: lim*cd + (1 - lim)*pos(faceFlux)
Every arithmetic operation is one fp64 rounding in the reference's order (numpy element-wise operations).
Pinned: tests/test_limiters_cpu.py runs the reference's own limiter headers (oracle/_ref/libref_limiters.so) beside it."""
import ctypes as C
import os

import numpy as np

SCHEMES = {"upwind": 0, "linear": 1, "limitedLinear": 2, "vanLeer": 3, "Minmod": 4}
SMALL = 1e-15


def nvdtvd_r(faceFlux, phiP, phiN, gradcP, gradcN, d):
    gradf = phiN - phiP
    g = np.where((faceFlux > 0)[:, None], gradcP, gradcN)
    gradcf = (d[:, 0] * g[:, 0] + d[:, 1] * g[:, 1]) + d[:, 2] * g[:, 2]
    sign = lambda x: np.where(x >= 0, 1.0, -1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        regular = 2.0 * (gradcf / gradf) - 1.0
    clipped = (2000.0 * sign(gradcf)) * sign(gradf) - 1.0
    return np.where(np.abs(gradcf) >= 1000.0 * np.abs(gradf), clipped, regular)


def limiter(scheme, lower, upper, faceFlux, vf, gradc, centres, k=1.0):
    lower, upper = np.asarray(lower), np.asarray(upper)
    nF = len(lower)
    if scheme == "upwind":
        return np.zeros(nF)
    if scheme == "linear":
        return np.ones(nF)
    d = centres[upper] - centres[lower]
    r = nvdtvd_r(np.asarray(faceFlux), vf[lower], vf[upper], gradc[lower], gradc[upper], d)
    if scheme == "limitedLinear":
        twoByk = 2.0 / max(k, SMALL)
        return np.maximum(np.minimum(twoByk * r, 1.0), 0.0)
    if scheme == "vanLeer":
        return (r + np.abs(r)) / (1.0 + np.abs(r))
    if scheme == "Minmod":
        return np.maximum(np.minimum(r, 1.0), 0.0)
    raise ValueError(f"Unknown discretisation scheme {scheme}")


def limited_weights(faceFlux, lim=None, cdWeights=None):
    p = np.where(np.asarray(faceFlux) >= 0, 1.0, 0.0)
    if lim is None:
        return p
    return lim * cdWeights + (1.0 - lim) * p


# ---- the reference's own limiter headers, compiled for the host (oracle/_ref/libref_limiters.so) ----
_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_limiters.so")


def reference_available():
    return os.path.exists(_LIB)


def reference_limiter(scheme, lower, upper, cdWeights, faceFlux, vf, gradc, centres, k=1.0):
    L = C.CDLL(_LIB)
    L.ref_limiter.argtypes = [C.c_int, C.c_double, C.c_int] + [C.c_void_p] * 8
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    f64 = lambda a: np.ascontiguousarray(a, np.float64)
    l, u, cd, fl, v, g, cc = i32(lower), i32(upper), f64(cdWeights), f64(faceFlux), f64(vf), f64(gradc), f64(centres)
    out = np.zeros(len(l))
    rc = L.ref_limiter(SCHEMES[scheme], float(k), len(l), *(a.ctypes.data for a in (l, u, cd, fl, v, g, cc, out)))
    if rc != 0:
        raise ValueError(f"ref_limiter: {rc}")
    return out
