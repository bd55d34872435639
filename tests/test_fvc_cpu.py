"""Host-side compositions of rapidcfd-dev_b200/fvc.py (explicit Laplacian, non-orthogonal snGrad correction) through the
oracle-backed stand-in, on a sheared (non-orthogonal, unskewed) hex mesh read through foamfile: for a linear field the
corrected face-normal gradient is exact and the explicit Laplacian vanishes."""
import importlib

import numpy as np

import oracle_backend


def test_corrected_sngrad_and_explicit_laplacian_on_a_sheared_mesh(meshmod, orc):
    run_sheared_case(meshmod, *oracle_backend.fixture())


def run_sheared_case(meshmod, capi, ctx, torch):
    """shared with the `-m gpu` test of tests/test_zzz_fvm_gpu.py (there capi is the CUDA library)"""
    ff = importlib.import_module("rapidcfd-dev_b200.foamfile")
    fvc = importlib.import_module("rapidcfd-dev_b200.fvc")
    m = meshmod.hex_mesh(6, 5, 4)
    pm = ff.from_hex_mesh(m)
    pm.points = pm.points + np.outer(pm.points[:, 1], [0.35, 0.0, 0.15])     # shear: x += 0.35 y, z += 0.15 y
    geo = pm.fv_geometry()
    nI = pm.nInternalFaces
    lower, upper = pm.ldu()
    bfc = pm.boundary_face_cells()
    C, Cf, Sf, V = geo["C"], geo["Cf"], geo["Sf"], geo["V"]
    magSf = geo["magSf"]
    n = Sf / magSf[:, None]
    d = C[upper] - C[lower]
    nonOrthDelta = 1.0 / np.maximum(np.einsum("ij,ij->i", n[:nI], d), 0.05 * np.linalg.norm(d, axis=1))   # surfaceInterpolation.C
    corrVecs = n[:nI] - d * nonOrthDelta[:, None]                                                         # :370-400
    assert np.abs(corrVecs).max() > 0.05                      # the mesh really is non-orthogonal
    g = np.array([0.7, -1.3, 0.4])
    psi = C @ g
    addr = capi.LduAddressing(ctx, pm.nCells, lower, upper)
    capi.fv_boundary_set(addr, bfc)
    ops = capi.FieldOps(ctx)
    t = lambda a: torch.from_numpy(np.array(a, dtype=np.float64).ravel()).to(ctx.device)
    h = lambda x: x.cpu().numpy()
    bvf = Cf[nI:] @ g                                         # exact boundary face values
    corr = fvc.snGrad_correction(capi, addr, ops, t(corrVecs), t(Sf[:nI]), t(geo["weights"]), t(psi), t(Sf[nI:]), t(bvf), t(V))
    sn = capi.fv_sngrad(addr, 1, t(nonOrthDelta), t(psi))
    np.testing.assert_allclose(h(sn) + h(corr), n[:nI] @ g, rtol=0, atol=1e-10)
    # explicit Laplacian of the linear field with exact boundary fluxes: zero
    gamma = 1.7
    bFlux = gamma * magSf[nI:] * (n[nI:] @ g)
    lap = fvc.laplacian(capi, addr, ops, 1, t(gamma * magSf[:nI]), t(nonOrthDelta), t(psi), t(bFlux), t(V), correction=corr)
    np.testing.assert_allclose(h(lap), 0, atol=1e-9)
    # without the correction the operator is what fvc::div(gamma*snGrad*magSf) gives with the uncorrected scheme
    lap0 = fvc.laplacian(capi, addr, ops, 1, t(gamma * magSf[:nI]), t(nonOrthDelta), t(psi), t(bFlux), t(V))
    flux = gamma * magSf[:nI] * (nonOrthDelta * (psi[upper] - psi[lower]))
    ref = np.zeros(pm.nCells)
    np.add.at(ref, lower, flux)
    np.subtract.at(ref, upper, flux)
    np.add.at(ref, bfc, bFlux)
    np.testing.assert_allclose(h(lap0), ref / V, rtol=1e-12, atol=1e-12)
    assert np.abs(h(lap0)).max() > 1e-3                  # ... which is not zero on this mesh
    addr.close()
