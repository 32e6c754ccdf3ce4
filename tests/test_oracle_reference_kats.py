"""Pins the CPU oracle against the known-answer tests the reference holds for the
implicit soil-column path.  Each test names the reference test it
transcribes (paths relative to /root/reference).  No GPU, no reference needed at
run time.  Tolerances are the reference's own (`≈` in Julia is rtol=sqrt(eps)).
"""
import math

import numpy as np
import pytest

import oracle as orc

L = orc.lib()
EPS = np.finfo(np.float64).eps
RTOL = math.sqrt(EPS)  # Julia isapprox default


def approx(a, b):
    return np.allclose(a, b, rtol=RTOL, atol=0.0)


# ----------------------------------------------------------------------------
# point functions: test/standalone/Soil/soil_parameterizations.jl
# ----------------------------------------------------------------------------
def test_brooks_corey_closure():
    """soil_parameterizations.jl:190-220"""
    psi_b, c = -0.09, 0.228
    S = [0.5, 1.0, 1.5]
    theta = [0.3, 0.4, 0.5]
    theta_r, nu, K_sat, S_s = 0.2, 0.4, 2.9e-7, 1e-2
    va = S[0] ** (-1 / c) * psi_b
    psi = [L.orc_bc_matric_potential(c, psi_b, s) for s in S[:2]]
    assert approx([L.orc_bc_inverse_matric_potential(c, psi_b, p) for p in psi], S[:2])
    assert approx(psi, [va, psi_b])
    p = [L.orc_bc_pressure_head(c, psi_b, theta_r, t, nu, S_s) for t in theta]
    assert approx(p, psi + [0.1 / 1e-2 + psi_b])
    dpsi = [L.orc_bc_dpsidtheta(c, psi_b, t, nu, theta_r, S_s) for t in theta]
    va = [-psi_b / (c * (nu - theta_r)) * s ** (-(1 + 1 / c)) for s in S]
    assert approx(dpsi, [va[0], 1 / S_s, 1 / S_s])
    k = [L.orc_bc_hydraulic_conductivity(c, K_sat, s) for s in S]
    assert approx(k, [S[0] ** (2 / c + 3) * K_sat, K_sat, K_sat])


def test_van_genuchten_closure():
    """soil_parameterizations.jl:222-281"""
    theta_r, nu, S_s, a, n, K_sat = 0.2, 0.4, 1e-2, 3.6, 1.56, 2.9e-7
    m = 1.0 - 1.0 / n
    theta = [0.3, 0.4, 0.5]
    S = [L.orc_effective_saturation(nu, t, theta_r) for t in theta]
    assert approx(S, [0.5, 1.0, 1.5])
    va = -((S[0] ** (-1 / m) - 1) * a ** (-n)) ** (1 / n)
    psi = [L.orc_vg_matric_potential(a, n, m, s) for s in S[:2]]
    assert approx([L.orc_vg_inverse_matric_potential(a, n, m, p) for p in psi], S[:2])
    assert approx(psi[0], va) and psi[1] == 0.0
    dpsi = [L.orc_vg_dpsidtheta(a, n, m, t, nu, theta_r, S_s) for t in theta]
    va = 1.0 / (a * m * n) / (nu - theta_r) * (S[0] ** (-1 / m) - 1) ** (1 / n - 1) * S[0] ** (-1 / m - 1)
    assert approx(dpsi, [va, 1 / S_s, 1 / S_s])
    p = [L.orc_vg_pressure_head(a, n, m, theta_r, t, nu, S_s) for t in theta]
    assert approx(p[0], psi[0]) and p[1] == 0.0 and approx(p[2], 10.0)
    k = [L.orc_vg_hydraulic_conductivity(m, K_sat, s) for s in S]
    va = (math.sqrt(S[0]) * (1 - (1 - S[0] ** (1 / m)) ** m) ** 2) * K_sat
    assert approx(k, [va, K_sat, K_sat])
    vlf = [L.orc_volumetric_liquid_fraction(t, 0.5, 0.0) for t in (0.25, 0.5, 0.75)]
    assert approx(vlf, [0.25, 0.5, 0.5])


def test_heat_point_functions():
    """soil_parameterizations.jl:104-186 (the four functions the implicit path uses)"""
    E = orc.EARTH
    rho_i, rho_l, cp_i, cp_l, T0, Lf = E["rho_i"], E["rho_l"], E["cp_i"], E["cp_l"], E["T_ref"], E["LH_f0"]
    assert L.orc_temperature_from_rho_e_int(5.4e7, 0.05, 2.1415e6, rho_i, T0, Lf) == \
        T0 + (5.4e7 + 0.05 * rho_i * Lf) / 2.1415e6
    rho_c_ds = 2e6 * (1 - 0.2)
    assert approx(L.orc_volumetric_heat_capacity(0.25, 0.05, rho_c_ds, rho_l, cp_l, rho_i, cp_i),
                  rho_c_ds + 0.25 * rho_l * cp_l + 0.05 * rho_i * cp_i)
    assert L.orc_volumetric_internal_energy(0.05, 2.1415e6, 300.0, rho_i, T0, Lf) == \
        2.1415e6 * (300.0 - T0) - 0.05 * rho_i * Lf
    assert L.orc_volumetric_internal_energy_liq(300.0, rho_l, cp_l, T0) == rho_l * cp_l * (300.0 - T0)
    assert approx(L.orc_impedance_factor(1.0, 7.0), 1e-7)
    for T in (278.0, 288.0, 298.0):
        assert approx(L.orc_viscosity_factor(T, 2.64e-2, 288.0), math.exp(2.64e-2 * (T - 288.0)))


def test_heaviside():
    """src/shared_utilities/utils.jl:83-99: x - a > eps, not >= 0"""
    assert L.orc_heaviside(0.0, 0.0) == 0.0
    assert L.orc_heaviside(EPS, 0.0) == 0.0
    assert L.orc_heaviside(2 * EPS, 0.0) == 1.0
    assert L.orc_heaviside(0.5, 0.4) == 1.0 and L.orc_heaviside(0.4, 0.5) == 0.0


# ----------------------------------------------------------------------------
# Jacobian entries: test/shared_utilities/implicit_timestepping/richards_model.jl
# ----------------------------------------------------------------------------
CLAY = dict(nu=0.495, K_sat=0.0443 / 3600 / 100, S_s=1e-3, hcm_b=1.43, hcm_a=2.6, theta_r=0.124)
CLAY["hcm_m"] = 1 - 1 / CLAY["hcm_b"]


def _clay_K_dpsi():
    K = L.orc_vg_hydraulic_conductivity(CLAY["hcm_m"], CLAY["K_sat"],
                                        L.orc_effective_saturation(CLAY["nu"], 0.24, CLAY["theta_r"]))
    d = L.orc_vg_dpsidtheta(CLAY["hcm_a"], CLAY["hcm_b"], CLAY["hcm_m"], 0.24, CLAY["nu"],
                            CLAY["theta_r"], CLAY["S_s"])
    return K, d


@pytest.mark.parametrize("ncol", [1, 3])  # Column and HybridBox(1,1,nelems) in the reference
def test_richards_jacobian_moisture_bc(ncol):
    """richards_model.jl:16-141: uniform theta=0.24, dz=0.01, dtgamma=1, top MoistureStateBC(nu-1e-3),
    bottom FreeDrainage"""
    N = 150
    P = orc.Problem(model=orc.RICHARDS, top_bc=orc.TOP_MOISTURE_STATE, bottom_bc=orc.BOT_FREE_DRAINAGE,
                    z_f=np.linspace(-1.5, 0.0, N + 1), ncol=ncol, theta_bc_top=CLAY["nu"] - 1e-3, **CLAY)
    Y, p, W = P.new_state(), P.new_cache(), P.new_jacobian()
    Y.theta_l[:] = 0.24
    P.update_implicit_cache(Y, p)          # uic!(p, Y, 0)
    P.compute_jacobian(W, Y, p, 1.0)
    K, d = _clay_K_dpsi()
    dz = 0.01
    for c in range(ncol):
        lo, di, up = W.w11_lo[c], W.w11_di[c], W.w11_up[c]
        assert approx(lo[1:], up[:-1])
        assert lo[0] == 0.0 and up[-1] == 0.0
        assert approx(lo[1:], 1.0 * (K / dz**2 * d))
        assert approx(di[0], 1.0 * (-K / dz**2 * d) - 1)
        assert approx(di[1:-1], 1.0 * (-2 * K / dz**2 * d) - 1)
        assert approx(di[-1], 1.0 * (-K / dz**2 * d - K / (dz * dz / 2) * d) - 1)


def test_richards_jacobian_flux_bc():
    """richards_model.jl:143-227: top WaterFluxBC(-K_sat), bottom FreeDrainage"""
    N = 150
    P = orc.Problem(model=orc.RICHARDS, top_bc=orc.TOP_FLUX, bottom_bc=orc.BOT_FREE_DRAINAGE,
                    z_f=np.linspace(-1.5, 0.0, N + 1), ncol=2, **CLAY)
    Y, p, W = P.new_state(), P.new_cache(), P.new_jacobian()
    Y.theta_l[:] = 0.24
    p.top_bc_w[:] = -CLAY["K_sat"]
    P.update_implicit_cache(Y, p)
    P.compute_jacobian(W, Y, p, 1.0)
    K, d = _clay_K_dpsi()
    dz = 0.01
    for c in range(2):
        di = W.w11_di[c]
        assert approx(di[0], (-K / dz**2 * d) - 1)
        assert approx(di[1:-1], (-2 * K / dz**2 * d) - 1)
        assert approx(di[-1], (-K / dz**2 * d) - 1)


def test_energy_hydrology_jacobian_flux_bc():
    """energy_hydrology_model.jl:16-173: theta=0.24, theta_i=0, T=280, zero flux BCs, dtgamma=1.
    K and kappa are lagged cache inputs for EnergyHydrology: the reference test builds K_ic with
    impedance*viscosity*vG K and kappa_ic from the Kersten model; here they are given numbers, and
    the entries are checked against the same formulas in K_ic, kappa_ic."""
    N = 150
    E = orc.EARTH
    Kvg, d = _clay_K_dpsi()
    K_ic = L.orc_impedance_factor(0.0, 7.0) * L.orc_viscosity_factor(280.0, 2.64e-2, 288.0) * Kvg
    kappa_ic = 1.37  # any positive number: the entries are linear in it
    rho_c_ds = 2.3e6 * (1 - CLAY["nu"])
    P = orc.Problem(model=orc.ENERGY_HYDROLOGY, z_f=np.linspace(-1.5, 0.0, N + 1), ncol=2,
                    rho_c_ds=rho_c_ds, K_lag=K_ic, kappa_lag=kappa_ic, theta_l_lag=0.24, **CLAY)
    Y, p, W = P.new_state(), P.new_cache(), P.new_jacobian()
    Y.theta_l[:] = 0.24
    Y.theta_i[:] = 0.0
    rho_c_s = L.orc_volumetric_heat_capacity(0.24, 0.0, rho_c_ds, E["rho_l"], E["cp_l"], E["rho_i"], E["cp_i"])
    Y.rho_e_int[:] = L.orc_volumetric_internal_energy(0.0, rho_c_s, 280.0, E["rho_i"], E["T_ref"], E["LH_f0"])
    P.update_implicit_cache(Y, p)
    assert approx(p.T, 280.0)
    P.compute_jacobian(W, Y, p, 1.0)
    dz = 0.01
    dTdrho = 1 / rho_c_s
    e_liq = L.orc_volumetric_internal_energy_liq(280.0, E["rho_l"], E["cp_l"], E["T_ref"])
    for c in range(2):
        di = W.w11_di[c]
        assert approx(di[0], (-K_ic / dz**2 * d) - 1)
        assert approx(di[1:-1], (-2 * K_ic / dz**2 * d) - 1)
        assert approx(di[-1], (-K_ic / dz**2 * d) - 1)
        di = W.w22_di[c]
        assert approx(di[0], (-kappa_ic / dz**2 * dTdrho) - 1)
        assert approx(di[1:-1], (-2 * kappa_ic / dz**2 * dTdrho) - 1)
        assert approx(di[-1], (-kappa_ic / dz**2 * dTdrho) - 1)
        di = W.w21_di[c]  # off-diagonal block, checked WITH the -I (energy_hydrology_model.jl:163-172)
        assert approx(di[0], (-e_liq * K_ic / dz**2 * d) - 1)
        assert approx(di[1:-1], (-2 * e_liq * K_ic / dz**2 * d) - 1)
        assert approx(di[-1], (-e_liq * K_ic / dz**2 * d) - 1)


# ----------------------------------------------------------------------------
# tendency: test/standalone/Soil/soiltest.jl
# ----------------------------------------------------------------------------
def test_richards_hydrostatic_zero_tendency():
    """soiltest.jl:15-90: hydrostatic profile above a water table at zmin => zero tendency, psi+z = -10"""
    N, zmin = 50, -10.0
    nu, n, a, theta_r = 0.495, 2.0, 2.6, 0.0
    m = 1 - 1 / n
    P = orc.Problem(model=orc.RICHARDS, z_f=np.linspace(zmin, 0.0, N + 1), ncol=1, nu=nu,
                    K_sat=0.0443 / 3600 / 100, S_s=1e-3, hcm_b=n, hcm_a=a, hcm_m=m, theta_r=theta_r)
    Y, p, dY = P.new_state(), P.new_cache(), P.new_state()
    S = (1 + (a * (P.z_c - zmin)) ** n) ** (-m)
    Y.theta_l[0] = S * (nu - theta_r) + theta_r
    P.update_implicit_cache(Y, p)
    P.compute_imp_tendency(dY, Y, p)
    assert np.mean(dY.theta_l) < EPS
    assert np.mean(p.psi[0] + P.z_c + 10.0) < 2 * EPS
    assert np.max(np.abs(dY.theta_l)) < 1e-15      # stronger than the reference asks


def test_energy_hydrology_tendency_matches_hand_built_flux_formula():
    """soiltest.jl:97-406: implicit tendency of theta_l and rho_e_int against the hand-built
    face-flux formula (arithmetic-mean face K, centre differences), tolerance 1e2*eps as there.
    theta(z) = nu/2 + nu/4 (z+0.5)^2 ... the reference uses an analytic profile; any smooth profile
    exercises the same stencil, so we use its dtheta/dz = nu/2 (z+0.5) form."""
    N, zmin = 200, -1.0
    nu, n, a, theta_r, S_s = 0.495, 2.0, 2.6, 0.1, 1e-3
    m = 1 - 1 / n
    K_sat = 0.0443 / 3600 / 100
    E = orc.EARTH
    z_f = np.linspace(zmin, 0.0, N + 1)
    z = 0.5 * (z_f[1:] + z_f[:-1])
    dz = 1.0 / N
    theta = nu / 2 + nu / 4 * (z + 0.5) ** 2
    T = 280.0 + 0.5 * (z + 0.5) ** 2 * 10
    Kc = np.array([L.orc_vg_hydraulic_conductivity(m, K_sat, L.orc_effective_saturation(nu, t, theta_r))
                   for t in theta])
    kappa = 1.0 + 0.3 * np.sin(3 * z)
    rho_c_ds = 2e6 * (1 - nu)
    P = orc.Problem(model=orc.ENERGY_HYDROLOGY, z_f=z_f, ncol=1, nu=nu, K_sat=K_sat, S_s=S_s, hcm_b=n,
                    hcm_a=a, hcm_m=m, theta_r=theta_r, rho_c_ds=rho_c_ds, K_lag=Kc[None, :],
                    kappa_lag=kappa[None, :], theta_l_lag=theta[None, :])
    Y, p, dY = P.new_state(), P.new_cache(), P.new_state()
    Y.theta_l[0] = theta
    rho_c_s = rho_c_ds + theta * E["rho_l"] * E["cp_l"]
    Y.rho_e_int[0] = rho_c_s * (T - E["T_ref"])
    P.update_implicit_cache(Y, p)
    P.compute_imp_tendency(dY, Y, p)
    # hand-built (numpy), as the reference test does
    psi = np.array([L.orc_vg_pressure_head(a, n, m, theta_r, t, nu, S_s) for t in theta])
    K_face = 0.5 * (Kc[1:] + Kc[:-1])
    h = psi + z
    flux = np.concatenate([[0.0], -K_face * (h[1:] - h[:-1]) / dz, [0.0]])
    expected = -(flux[1:] - flux[:-1]) / dz
    assert np.mean(np.abs(expected - dY.theta_l[0])) / nu < 1e2 * EPS
    Tc = p.T[0]
    e_l = E["rho_l"] * E["cp_l"] * (Tc - E["T_ref"])
    eK_face = 0.5 * ((e_l * Kc)[1:] + (e_l * Kc)[:-1])
    k_face = 0.5 * (kappa[1:] + kappa[:-1])
    flux = np.concatenate([[0.0], -k_face * (Tc[1:] - Tc[:-1]) / dz - eK_face * (h[1:] - h[:-1]) / dz, [0.0]])
    expected = -(flux[1:] - flux[:-1]) / dz
    assert np.mean(np.abs(expected - dY.rho_e_int[0])) / np.median(Y.rho_e_int[0]) < 1e2 * EPS
    assert np.all(dY.theta_i == 0.0)


def test_flux_bc_conservation_signs():
    """conservation.jl:103-151 and :218-258: d(intF)/dt = -(F_top - F_bot) = -2 for F_top=1, F_bot=-1;
    column-integrated tendency equals the same."""
    N = 20
    for model in (orc.RICHARDS, orc.ENERGY_HYDROLOGY):
        P = orc.Problem(model=model, z_f=np.linspace(-1.0, 0.0, N + 1), ncol=2, nu=0.495,
                        K_sat=0.0443 / 3600 / 100, S_s=1e-3, hcm_b=2.0, hcm_a=2.6, hcm_m=0.5, theta_r=0.0,
                        rho_c_ds=1e6, K_lag=1e-7, kappa_lag=1.5, theta_l_lag=0.2475)
        Y, p, dY = P.new_state(), P.new_cache(), P.new_state()
        Y.theta_l[:] = 0.495 / 2
        Y.rho_e_int[:] = 2.0e7
        p.top_bc_w[:], p.bot_bc_w[:] = 1.0, -1.0
        p.top_bc_h[:], p.bot_bc_h[:] = 1.0, -1.0
        P.update_implicit_cache(Y, p)
        P.compute_imp_tendency(dY, Y, p)
        assert np.all(dY.intF_w == -2.0)
        assert np.allclose(P.column_integral(dY.theta_l), -2.0, rtol=1e-12)
        if model == orc.ENERGY_HYDROLOGY:
            assert np.all(dY.intF_e == -2.0)
            assert np.allclose(P.column_integral(dY.rho_e_int), -2.0, rtol=1e-9)


def test_moisture_state_bc_flux():
    """soil_bc.jl:98-133: state -> flux conversion, diffusive_flux(K_c, psi_bc + dz, psi_c, dz)"""
    N = 50
    nu, n, a = 0.495, 2.0, 2.6
    m = 0.5
    K_sat, S_s = 0.0443 / 3600 / 100, 1e-3
    P = orc.Problem(model=orc.RICHARDS, top_bc=orc.TOP_MOISTURE_STATE, bottom_bc=orc.BOT_MOISTURE_STATE,
                    z_f=np.linspace(-10.0, 0.0, N + 1), ncol=1, nu=nu, K_sat=K_sat, S_s=S_s, hcm_b=n, hcm_a=a,
                    hcm_m=m, theta_r=0.0, theta_bc_top=nu / 2, theta_bc_bot=nu / 2)
    Y, p = P.new_state(), P.new_cache()
    Y.theta_l[:] = nu / 3
    P.update_implicit_cache(Y, p)
    dz = 10.0 / N / 2.0
    K_c = L.orc_vg_hydraulic_conductivity(m, K_sat, L.orc_effective_saturation(nu, nu / 3, 0.0))
    psi_bc = L.orc_vg_pressure_head(a, n, m, 0.0, nu / 2, nu, S_s)
    psi_c = L.orc_vg_pressure_head(a, n, m, 0.0, nu / 3, nu, S_s)
    assert abs(p.top_bc_w[0] - (-K_c * ((psi_bc - psi_c + dz) / dz))) < 1e-20
    assert abs(p.bot_bc_w[0] - (-K_c * ((psi_c + dz - psi_bc) / dz))) < 1e-20
    assert approx(p.dfluxBCdY[0], K_c * L.orc_vg_dpsidtheta(a, n, m, nu / 3, nu, 0.0, S_s) / dz)


def test_free_drainage_and_total_water():
    """boundary_conditions.jl:340-353 (bottom_bc = -K_1) and conservation.jl:127-133
    (total water = nu/2 * depth)"""
    N = 30
    P = orc.Problem(model=orc.RICHARDS, top_bc=orc.TOP_MOISTURE_STATE, bottom_bc=orc.BOT_FREE_DRAINAGE,
                    z_f=-np.geomspace(1.0, 11.0, N + 1)[::-1] + 1.0, ncol=2, theta_bc_top=0.4, **CLAY)
    Y, p = P.new_state(), P.new_cache()
    Y.theta_l[:] = CLAY["nu"] / 2
    P.update_implicit_cache(Y, p)
    assert np.all(p.bot_bc_w == -p.K[:, 0])
    assert np.allclose(p.total_water, CLAY["nu"] / 2 * 10.0, atol=RTOL)
