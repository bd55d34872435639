"""one icoFoam step restated (oracle/piso_oracle.py) has to produce the physics of
the lid-driven cavity -- there is no reference build to compare with, so these are property checks."""
import numpy as np

from oracle import piso_oracle as po


def test_cavity_steps_conserve_mass_and_spin_up(meshmod, orc):
    n = 8
    m, case = po.cavity_from_hex(orc, meshmod, n, nu=0.01)
    ke, conts = [], []
    for step in range(6):
        perfs, cont = case.step(nCorr=2, pControls=dict(tolerance=1e-10, relTol=0.0),
                                UControls=dict(tolerance=1e-10, relTol=0.0))
        assert all(p.converged for p in perfs["U"]) and all(p.converged for p in perfs["p"])
        assert len(perfs["p"]) == 2                               # one pressure solve per PISO corrector
        conts.append(cont[-1][0])
        ke.append(0.5 * (case.U ** 2).sum(axis=1) @ case.V)
    # the corrected face flux is divergence free to the pressure solver's tolerance
    assert max(conts) < 1e-9
    div = case.div(case.phi, case.bphi)
    assert np.abs(div).max() < 1e-6
    # the lid drags the fluid along: kinetic energy grows monotonically during spin-up, the top layer moves with
    # the lid, the return flow below goes the other way, nothing leaves through the walls
    assert all(b > a for a, b in zip(ke, ke[1:]))
    cc = m.cell_centres()
    top = cc[:, 1] > 1 - m.h
    assert case.U[top, 0].mean() > 0.1
    assert case.U[(cc[:, 1] > 0.4) & (cc[:, 1] < 0.7), 0].mean() < 0
    assert np.array_equal(case.bphi, np.zeros_like(case.bphi))
    # the case is symmetric about the mid-plane z = 1/2: so is the solution (to solver tolerance)
    idx = np.arange(m.nCells).reshape(n, n, n)                    # [k, j, i]
    mirror = idx[::-1].ravel()
    np.testing.assert_allclose(case.U[mirror, 0], case.U[:, 0], atol=1e-8)
    np.testing.assert_allclose(case.U[mirror, 2], -case.U[:, 2], atol=1e-8)
    np.testing.assert_allclose(case.p[mirror], case.p, atol=1e-8)


def test_pressure_correction_projects_the_flux(meshmod, orc):
    """One corrector with an already converged momentum field: phi = phiHbyA - flux(p) removes exactly the divergence
    that div(phiHbyA) had put into the pressure equation's source."""
    m, case = po.cavity_from_hex(orc, meshmod, 6, nu=0.05)
    for _ in range(3):
        case.step(nCorr=2, pControls=dict(tolerance=1e-12, relTol=0.0), UControls=dict(tolerance=1e-12, relTol=0.0))
    U_before = case.U.copy()
    perfs, cont = case.step(nCorr=3, pControls=dict(tolerance=1e-12, relTol=0.0),
                            UControls=dict(tolerance=1e-12, relTol=0.0))
    assert cont[-1][0] < 1e-11 and abs(cont[-1][1]) < 1e-12
    # successive correctors converge: the later ones need fewer pressure iterations than the first
    its = [p.nIterations for p in perfs["p"]]
    assert its[-1] <= its[0]
    assert np.abs(case.U - U_before).max() < 0.2                  # a time step changes the field smoothly


def test_gamg_pressure_solver_gives_the_same_step(meshmod, orc):
    """fvSolution `p { solver GAMG; smoother GaussSeidel; }`: the step is the same to the solvers' tolerance"""
    n = 8
    m, a = po.cavity_from_hex(orc, meshmod, n)
    _, b = po.cavity_from_hex(orc, meshmod, n)
    g = orc.Gamg(b.addr, meshmod.face_area_pair_weights(m), 10)
    ctl = dict(tolerance=1e-12, relTol=0.0)
    for _ in range(2):
        a.step(UControls=ctl, pControls=ctl)
        perfs, _ = b.step(UControls=ctl, pControls=dict(ctl, maxIter=200), pSolver=("GAMG", "GaussSeidel"), gamg=g)
        assert all(p.converged for p in perfs["p"]) and perfs["p"][0].solverName == b"GAMG"
    np.testing.assert_allclose(b.U, a.U, atol=1e-9)
    np.testing.assert_allclose(b.p, a.p, atol=1e-9)


import pytest  # noqa: E402

import dist_helpers as dh  # noqa: E402


@pytest.mark.parametrize("nR", [2, 4, 8])
def test_decomposed_cavity_reproduces_the_single_domain(meshmod, orc, nR):
    """icoFoam over processor patches: every rank runs the same step on its brick (interpolation, fluxes, matrix
    coefficients and the glue on the coupled faces, halo exchanges, global sums) and the fields agree with the
    single-domain run to the solvers' tolerance."""
    n = 8
    ctl = dict(tolerance=1e-12, relTol=0.0)
    _, ref = po.cavity_from_hex(orc, meshmod, n)
    for _ in range(3):
        rp, rc = ref.step(UControls=ctl, pControls=ctl)
    ex = dh.ThreadExchange(nR)

    def rank_fn(r):
        m0 = meshmod.decompose(n, nR, r)
        comm = ex.comm(orc, r, m0, n ** 3)
        m, case = po.cavity_rank(orc, meshmod, n, nR, r, comm)
        for _ in range(3):
            perfs, cont = case.step(UControls=ctl, pControls=ctl)
        return m.cellGlobal, case.U, case.p, cont, [p.converged for p in perfs["U"] + perfs["p"]], case
    res = dh.run_threads(nR, rank_fn)
    for cg, U, p, cont, conv, case in res:
        assert all(conv)
        np.testing.assert_allclose(U, ref.U[cg], rtol=0, atol=1e-8)
        np.testing.assert_allclose(p, ref.p[cg], rtol=0, atol=1e-8)
        assert cont[-1][0] < 1e-10 and cont == res[0][3]          # global sums: the same numbers on every rank
    # both sides of a processor patch hold the same flux with opposite sign
    for r, (cg, U, p, cont, conv, case) in enumerate(res):
        ps = case.addr.patch_start()
        for i, nb in enumerate(case.addr.neighbRank):
            other = res[nb][5]
            q = list(other.addr.neighbRank).index(r)
            qs = other.addr.patch_start()
            np.testing.assert_allclose(case.cphi[ps[i]:ps[i + 1]], -other.cphi[qs[q]:qs[q + 1]], rtol=0, atol=1e-14)


def test_simple_iterations_converge_to_the_steady_cavity(meshmod, orc):
    """SIMPLE (simpleFoam/UEqn.H, pEqn.H; laminar, upwind convection, under-relaxation 0.7 / 0.3): the residuals of both
    equations fall over the iterations, continuity is enforced to the pressure solver's tolerance every iteration, and the
    steady field is the same recirculation the PISO run spins up to (top layer follows the lid, return flow below)."""
    n = 8
    m, c = po.cavity_from_hex(orc, meshmod, n, nu=0.05)
    ctl = dict(tolerance=1e-10, relTol=0.0)
    first, last = None, None
    for it in range(60):
        perfs, cont = c.simple_step(UControls=ctl, pControls=ctl)
        assert abs(cont[1]) < 1e-9                       # global continuity after every pressure correction
        last = (perfs["U"][0].initialResidual, perfs["p"][0].initialResidual)
        first = first or last
    assert last[0] < 1e-3 * first[0] and last[1] < 1e-2 * first[1]
    cc = m.cell_centres()
    top = c.U[cc[:, 1] > 1 - m.h, 0].mean()
    bottom = c.U[cc[:, 1] < 0.4, 0].mean()
    assert top > 0.1 and bottom < 0                      # lid-driven recirculation
    # linear convection gives another (close) steady state through the same path
    m2, c2 = po.cavity_from_hex(orc, meshmod, n, nu=0.05)
    for it in range(60):
        c2.simple_step(divScheme="linear", UControls=ctl, pControls=ctl)
    assert np.abs(c2.U - c.U).max() < 0.15 and np.abs(c2.U - c.U).max() > 1e-6
