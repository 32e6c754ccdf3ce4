"""SoilCO2Model's implicit diffusion in the CPU oracle.  The reference pins this path only by
properties (test/standalone/Soil/Biogeochemistry/biogeochemistry_module.jl:107-203: with zero-flux boundaries the
tendency integrates to zero and everything stays finite); they are transcribed here, together with the structural
identities that tie the block to the Richards Jacobian the reference does pin."""
import numpy as np

import oracle as orc


def _setup(ncol=5, N=20, seed=0, atm=False):
    rng = np.random.default_rng(seed)
    z_f = -np.cumsum(np.concatenate([[0.0], rng.uniform(0.02, 0.2, N)]))[::-1].copy()
    P = orc.Problem(model=orc.RICHARDS, z_f=z_f, ncol=ncol, nu=0.5, theta_r=0.1, K_sat=1e-6, S_s=1e-3, hcm_a=2.0,
                    hcm_b=2.0, hcm_m=0.5)
    D = rng.uniform(1e-8, 2e-6, (ncol, N))
    th = rng.uniform(0.02, 0.45, (ncol, N))
    S = P.co2_species(D, th, c_atm=rng.uniform(1e-4, 4e-4, ncol) if atm else None)
    C = rng.uniform(5e-5, 2e-3, (ncol, N))
    return P, S, C, z_f


def test_zero_flux_tendency_conserves_mass():
    """biogeochemistry_module.jl:199: sum(dY.soilco2.CO2) ~ 0 (there on a uniform grid; here dz-weighted)"""
    P, S, C, z_f = _setup()
    zero = np.zeros(P.ncol)
    dC = P.co2_imp_tendency(S, C, zero, zero)
    assert np.all(np.isfinite(dC))
    assert np.all(np.abs((dC * np.diff(z_f)).sum(axis=1)) <= 1e-15 * np.abs(dC * np.diff(z_f)).sum(axis=1))


def test_boundary_fluxes_enter_with_the_reference_sign():
    """flux BCs: d(sum C dz)/dt = -(top - bottom) (the soil water convention, conservation.jl:139-151)"""
    P, S, C, z_f = _setup(seed=1)
    top, bot = np.full(P.ncol, 3e-9), np.full(P.ncol, -1e-9)
    dC = P.co2_imp_tendency(S, C, top, bot)
    assert np.allclose((dC * np.diff(z_f)).sum(axis=1), -(top - bot), rtol=1e-12)


def test_jacobian_is_the_derivative_of_the_tendency():
    """W = dtgamma dT/dC - I for C > 0 (finite differences on the piecewise-linear tendency are exact up to
    rounding), including the dfluxBCdY entry of the state boundary condition (:1152-1166)"""
    P, S, C, z_f = _setup(ncol=2, N=9, seed=2, atm=True)
    top, dfl = np.zeros(P.ncol), np.zeros(P.ncol)
    P.co2_boundary_flux(S, C, top, dfl)
    bot = np.zeros(P.ncol)
    dtg = 700.0
    lo, di, up = P.co2_jacobian(S, dtg, dfl)
    N = C.shape[1]
    J = np.zeros((P.ncol, N, N))
    base = P.co2_imp_tendency(S, C, top, bot)
    for j in range(N):
        Cp = C.copy()
        h = 1e-6 * C[:, j]
        Cp[:, j] += h
        tp = top.copy()
        P.co2_boundary_flux(S, Cp, tp, np.zeros(P.ncol))
        J[:, :, j] = (P.co2_imp_tendency(S, Cp, tp, bot) - base) / h[:, None]
    W = dtg * J - np.eye(N)
    for c in range(P.ncol):
        assert np.allclose(np.diag(W[c]), di[c], rtol=1e-6)
        assert np.allclose(np.diag(W[c], -1), lo[c, 1:], rtol=1e-6)
        assert np.allclose(np.diag(W[c], 1), up[c, :-1], rtol=1e-6)
        assert np.allclose(np.triu(W[c], 2), 0.0, atol=1e-9 * np.abs(di[c]).max())


def test_linear_problem_converges_in_one_newton_iteration_and_solves_backward_euler():
    """for C > 0 the stage is linear: the first Newton update lands on the backward-Euler solution
    (I - dtgamma A) C1 = C0 + dtgamma b; later iterations change nothing beyond rounding"""
    import scipy.linalg
    P, S, C, z_f = _setup(ncol=3, N=15, seed=3)
    top, bot = np.full(P.ncol, 2e-10), np.zeros(P.ncol)
    dtg = 1800.0
    C1, C3 = C.copy(), C.copy()
    P.co2_implicit_step(S, C1, top.copy(), bot, dtg, 1)
    P.co2_implicit_step(S, C3, top.copy(), bot, dtg, 3)
    assert np.allclose(C1, C3, rtol=1e-13)
    lo, di, up = P.co2_jacobian(S, dtg)
    b0 = P.co2_imp_tendency(S, np.zeros_like(C), top, bot)  # the boundary-flux part of the tendency
    for c in range(P.ncol):
        ab = np.zeros((3, C.shape[1]))
        ab[0, 1:], ab[1], ab[2, :-1] = up[c, :-1], di[c], lo[c, 1:]
        want = scipy.linalg.solve_banded((1, 1), -ab, C[c] + dtg * b0[c])
        assert np.allclose(C1[c], want, rtol=1e-12)
    # mass balance of the stage against the boundary fluxes
    dz = np.diff(z_f)
    assert np.allclose(((C3 - C) * dz).sum(axis=1), -dtg * (top - bot), rtol=1e-10)
