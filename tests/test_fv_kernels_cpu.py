"""The cell-parallel face-sum kernels of csrc/fv.cu (fv_kernels.cuh: surfaceIntegrate / surfaceSum, gaussGrad::gradf, the fused
interpolate + gradient, negSumDiag) executed on the host and compared bit for bit with the oracle -- on the hex mesh (three
faces per side: the prefetched batch) and on a random graph (up to a dozen faces per side: the loops after the batch, cells
without owner or neighbour faces)."""
import ctypes as C

import numpy as np
import pytest

from test_host_kernels_cpu import Host, _d, hk, run_all  # noqa: F401  (hk: fixture, the host build of the kernels)


class _Graph:
    def __init__(self, n, k, seed):
        rng = np.random.default_rng(seed)
        a, b = rng.integers(0, n, size=n * k), rng.integers(0, n, size=n * k)
        keep = a != b
        pr = np.unique(np.stack([np.minimum(a, b)[keep], np.maximum(a, b)[keep]], 1), axis=0).astype(np.int32)
        self.lower, self.upper = pr[:, 0].copy(), pr[:, 1].copy()
        self.nCells, self.nFaces = n, len(pr)


def _cases(meshmod):
    m = meshmod.hex_mesh(7, 6, 5)
    _, bfc = m.patch_start_facecells(m.wall_patches())
    yield "hex", m, bfc
    g = _Graph(300, 4, 2)
    yield "graph", g, np.random.default_rng(3).integers(0, g.nCells, 170).astype(np.int32)   # several boundary faces per cell
    yield "graph-no-boundary", _Graph(200, 3, 4), np.zeros(0, np.int32)


@pytest.mark.parametrize("nc", [1, 3])
def test_face_sum_kernels_on_the_host(hk, meshmod, orc, nc):  # noqa: F811
    for name, m, bfc in _cases(meshmod):
        rng = np.random.default_rng(7)
        n, nF, nB = m.nCells, m.nFaces, len(bfc)
        a = orc.Addr(n, m.lower, m.upper)
        H = Host(a, dict(bfc=bfc, diag=np.zeros(n), upper=np.zeros(nF), lower=None))
        f = lambda x: np.ascontiguousarray(x, np.float64)
        ssf, bssf, V = f(rng.uniform(-1, 1, (nF, nc))), f(rng.uniform(-1, 1, (max(nB, 1), nc))), f(rng.uniform(0.5, 2, n))
        Sf, bSf = f(rng.uniform(-1, 1, (nF, 3))), f(rng.uniform(-1, 1, (max(nB, 1), 3)))
        w, vf = f(rng.uniform(0, 1, nF)), f(rng.uniform(-1, 1, (n, nc)))
        hk.hk_surface_integrate.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int]
        for div, sign in ((1, -1), (0, 1), (0, -1)):
            out = np.zeros((n, nc))
            hk.hk_surface_integrate(H.p(), nc, _d(ssf), _d(bssf), _d(V), _d(out), div, sign)
            want = np.asarray(orc.surface_integrate(a, ssf.ravel(), bfc, bssf[:nB].ravel(), V, nc, bool(div), sign)).reshape(n, nc)
            assert np.array_equal(out, want), (name, div, sign)
        hk.hk_gauss_grad.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        out = np.zeros((n, 3 * nc))
        hk.hk_gauss_grad(H.p(), nc, _d(Sf), _d(ssf), _d(bSf), _d(bssf), _d(V), _d(out))
        want = np.asarray(orc.gauss_grad(a, Sf.ravel(), ssf.ravel(), bfc, bSf[:nB].ravel(), bssf[:nB].ravel(), V, nc)).reshape(n, 3 * nc)
        assert np.array_equal(out, want), name
        hk.hk_grad_linear.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
        out = np.zeros((n, 3 * nc))
        hk.hk_grad_linear(H.p(), nc, _d(Sf), _d(w), _d(vf), _d(bSf), _d(bssf), _d(V), _d(out))
        face = np.asarray(orc.interpolate_linear(a, w, vf.ravel(), nc))
        want = np.asarray(orc.gauss_grad(a, Sf.ravel(), face.ravel(), bfc, bSf[:nB].ravel(), bssf[:nB].ravel(), V, nc)).reshape(n, 3 * nc)
        assert np.array_equal(out, want), name


def test_neg_sum_diag_on_the_host(hk, meshmod, orc):  # noqa: F811
    for name, m, bfc in _cases(meshmod):
        rng = np.random.default_rng(9)
        n, nF = m.nCells, m.nFaces
        a = orc.Addr(n, m.lower, m.upper)
        H = Host(a, dict(bfc=bfc, diag=np.zeros(n), upper=np.zeros(nF), lower=None))
        delta, g, wts, phi = (rng.uniform(0.1, 1, nF) for _ in range(4))
        up, dg = (np.asarray(x) for x in orc.laplacian_fill(a, delta, g))
        hk.hk_neg_sum_diag.argtypes = [C.c_void_p] * 4
        out = np.zeros(n)
        hk.hk_neg_sum_diag(H.p(), _d(np.ascontiguousarray(up)), _d(np.ascontiguousarray(up)), _d(out))
        assert np.array_equal(out, dg), name
        lo, up2, dg2 = (np.ascontiguousarray(x) for x in orc.convection_fill(a, wts, phi - 0.5))
        hk.hk_neg_sum_diag(H.p(), _d(up2), _d(lo), _d(out))
        assert np.array_equal(out, dg2), name


@pytest.mark.parametrize("nc", [1, 3])
def test_fvmatrix_kernels_on_a_random_graph(hk, orc, nc):  # noqa: F811
    """A, H, flux, boundary folding, relax, setReference (csrc/fvmatrix_kernels.cuh) on rows with up to a dozen faces per side and
    several boundary faces per cell: the loops behind the prefetched first batch of H and relax"""
    g = _Graph(260, 4, 11)
    rng = np.random.default_rng(12)
    n, nF = g.nCells, g.nFaces
    bfc = rng.integers(0, n, 140).astype(np.int32)
    a = orc.Addr(n, g.lower, g.upper)
    upper, lower = rng.uniform(-1, 1, nF), rng.uniform(-1, 1, nF)
    d = dict(diag=rng.uniform(8, 12, n), upper=upper, lower=None if nc == 1 else lower, source=rng.uniform(-1, 1, (n, nc)) if nc > 1 else
             rng.uniform(-1, 1, n), bfc=bfc, ic=rng.uniform(0, 1, (len(bfc), nc)), bc=rng.uniform(-1, 1, (len(bfc), nc)),
             V=rng.uniform(0.5, 2, n))
    assert np.bincount(g.lower, minlength=n).max() > 3 and np.bincount(g.upper, minlength=n).max() > 3
    run_all(hk, orc, a, d, nc, rng.uniform(-1, 1, (n, nc)))
