"""Pins the CPU oracle's explicit-stage functions (update_aux!, PhaseChange) against
the known-answer tests the reference holds for them.  Each test names the reference test it transcribes
(paths relative to /root/reference).  No GPU.  `≈` in Julia is rtol = sqrt(eps)."""
import math

import numpy as np

import oracle as orc

L = orc.lib()
EPS = np.finfo(np.float64).eps
RTOL = math.sqrt(EPS)
E = orc.EARTH
X = orc.EXPLICIT_SCALARS


def approx(a, b, atol=0.0):
    return np.allclose(a, b, rtol=RTOL, atol=atol)


def test_heat_parameterizations():
    """test/standalone/Soil/soil_parameterizations.jl:122-156 (kappa_sat, relative_saturation,
    kersten_number with and without ice, thermal_conductivity)"""
    assert approx(L.orc_kappa_sat(0.25, 0.05, 0.57, 2.29), 0.57 ** (0.25 / (0.05 + 0.25)) * 2.29 ** (0.05 / (0.05 + 0.25)))
    assert L.orc_kappa_sat(0.0, 0.0, 0.57, 2.29) == 0.5 * (0.57 + 2.29)
    assert L.orc_relative_saturation(0.25, 0.05, 0.4) == (0.25 + 0.05) / 0.4
    om = quartz = gravel = 0.1
    want = 0.75 ** ((1 + 0.1 - 0.24 * 0.1 - 0.1) / 2) * ((1 + math.exp(-18.3 * 0.75)) ** (-3) - ((1 - 0.75) / 2) ** 3) ** (1 - 0.1)
    assert approx(L.orc_kersten_number(0.0, 0.75, X["alpha"], X["beta"], om, quartz, gravel), want)
    assert L.orc_kersten_number(0.05, 0.75, X["alpha"], X["beta"], om, quartz, gravel) == 0.75 ** (1 + 0.1)
    assert L.orc_thermal_conductivity(1.5, 0.7287, 0.7187) == 0.7287 * 0.7187 + (1 - 0.7287) * 1.5


def test_impedance_and_viscosity():
    """soil_parameterizations.jl:178-185"""
    assert approx(L.orc_impedance_factor(1.0, X["Omega"]), 1e-7)
    T = np.array([278.0, 288.0, 298.0])
    got = [L.orc_viscosity_factor(t, X["gamma"], X["gammaT_ref"]) for t in T]
    assert approx(got, np.exp(X["gamma"] * (T - X["gammaT_ref"])))


def _phase_change_expected(theta_l, theta_i, T, nu, theta_r, a, n, m, rho_c_ds, kappa_dry, dz):
    """the reference test's own re-derivation (soil_parameterizations.jl:329-341)"""
    theta_tot = E["rho_i"] / E["rho_l"] * theta_i + theta_l
    psi0 = L.orc_vg_matric_potential(a, n, m, L.orc_effective_saturation(nu, theta_tot, theta_r))
    Tf = X["T_freeze"] * math.exp(X["grav"] * psi0 / E["LH_f0"])
    psi_T = E["LH_f0"] / X["grav"] * math.log(T / Tf) * L.orc_heaviside(Tf - T, 0.0)
    theta_star = L.orc_vg_inverse_matric_potential(a, n, m, psi0 + psi_T) * (nu - theta_r) + theta_r
    rho_c = L.orc_volumetric_heat_capacity(theta_l, theta_i, rho_c_ds, E["rho_l"], E["cp_l"], E["rho_i"], E["cp_i"])
    tau = L.orc_thermal_time(rho_c, dz, kappa_dry)
    return theta_star, tau


def _src(theta_l, theta_i, T, tau, nu, theta_r, a, n, m):
    return L.orc_phase_change_source(orc.VAN_GENUCHTEN, a, n, m, theta_l, theta_i, T, tau, nu, theta_r, E["rho_i"],
                                     E["rho_l"], E["LH_f0"], X["T_freeze"], X["grav"])


def test_freezing_and_thawing():
    """soil_parameterizations.jl:283-430 ("Freezing and Thawing"): the source equals
    (theta_l - theta_star) / tau, is positive when freezing (T = 270 K, no ice), zero above the depressed
    freezing point without ice, and negative (melting) above it with ice."""
    nu, theta_r, a, n = 0.2, 0.1, 2.0, 1.4
    m = 1.0 - 1.0 / n
    rho_c_ds, kappa_dry, dz = 1.8e6, 1.1, 1.0   # any positive values: the test re-derives tau from them
    assert L.orc_thermal_time(rho_c_ds, dz, kappa_dry) == 3 * rho_c_ds * dz ** 2 / kappa_dry
    # freezing: T = 270 K
    got, want = [], []
    for theta_l in (0.11, 0.15, nu):
        ts, tau = _phase_change_expected(theta_l, 0.0, 270.0, nu, theta_r, a, n, m, rho_c_ds, kappa_dry, dz)
        got.append(_src(theta_l, 0.0, 270.0, tau, nu, theta_r, a, n, m))
        want.append((theta_l - ts) / tau)
    assert approx(got, want) and all(g > 0.0 for g in got)
    # T = 274 K, no ice: nothing happens, theta_star == theta_l
    for theta_l in (0.11, 0.15, nu):
        ts, tau = _phase_change_expected(theta_l, 0.0, 274.0, nu, theta_r, a, n, m, rho_c_ds, kappa_dry, dz)
        assert approx(ts, theta_l)
        assert abs(_src(theta_l, 0.0, 274.0, tau, nu, theta_r, a, n, m)) <= RTOL * theta_l / tau
    # T = 274 K with ice: melting
    got, want = [], []
    for theta_i in (0.05, 0.08):
        ts, tau = _phase_change_expected(0.11, theta_i, 274.0, nu, theta_r, a, n, m, rho_c_ds, kappa_dry, dz)
        got.append(_src(0.11, theta_i, 274.0, tau, nu, theta_r, a, n, m))
        want.append((0.11 - ts) / tau)
    assert approx(got, want) and all(g < 0.0 for g in got)


def _problem(ncol=40, N=15, seed=3, closure=0):
    import climaland_b200  # noqa: F401
    from climaland_b200 import workloads
    from helpers import oracle_problem
    w = workloads.make_workload("energy_hydrology", ncol, N=N, seed=seed, topmodel=False)
    xp = workloads.make_explicit_params(w, seed)
    P, Y, p = oracle_problem(w, closure=closure)
    return w, xp, P, Y


def test_update_aux_against_point_functions():
    """update_aux! (energy_hydrology.jl:722-814) assembled from the pinned point functions, cell by
    cell, including the two argument orders of soil_Tf_depressed the reference uses (:800-811 passes
    (rho_l, rho_i) where the function expects (rho_ice, rho_liq))."""
    w, xp, P, Y = _problem()
    Xp = P.explicit_params(**xp)
    a = P.new_aux()
    P.update_aux(Xp, Y, a)
    dz = np.diff(w["z_f"])
    for c in (0, 7, 39):
        for i in (0, 6, 14):
            nu, thr = w["nu"][c, i], w["theta_r"][c, i]
            th, thi, rho_e = Y.theta_l[c, i], Y.theta_i[c, i], Y.rho_e_int[c, i]
            al, n, m = w["hcm_a"][c, i], w["hcm_b"][c, i], w["hcm_m"][c, i]
            tl = L.orc_volumetric_liquid_fraction(th, nu - thi, thr)
            assert a.theta_l[c, i] == tl
            Ke = L.orc_kersten_number(thi, (tl + thi) / nu, X["alpha"], X["beta"], xp["nu_ss_om"][c, i],
                                      xp["nu_ss_quartz"][c, i], xp["nu_ss_gravel"][c, i])
            ks = L.orc_kappa_sat(tl, thi, xp["kappa_sat_unfrozen"][c, i], xp["kappa_sat_frozen"][c, i])
            assert a.kappa[c, i] == L.orc_thermal_conductivity(xp["kappa_dry"][c, i], Ke, ks)
            rc = L.orc_volumetric_heat_capacity(tl, thi, w["rho_c_ds"][c, i], E["rho_l"], E["cp_l"], E["rho_i"], E["cp_i"])
            T = L.orc_temperature_from_rho_e_int(rho_e, thi, rc, E["rho_i"], E["T_ref"], E["LH_f0"])
            assert a.T[c, i] == T
            K = (L.orc_impedance_factor(thi / (tl + thi - thr), X["Omega"]) * L.orc_viscosity_factor(T, X["gamma"], X["gammaT_ref"])
                 * L.orc_vg_hydraulic_conductivity(m, w["K_sat"][c, i], L.orc_effective_saturation(nu, th, thr)))
            assert a.K[c, i] == K
            assert a.psi[c, i] == L.orc_vg_pressure_head(al, n, m, thr, th, nu - thi, w["S_s"][c, i])
            assert a.Tf_depressed[c, i] == L.orc_soil_Tf_depressed(0, al, n, m, tl, thi, nu, thr, E["rho_l"], E["rho_i"],
                                                                   X["T_freeze"], X["grav"], E["LH_f0"])
    # column integrals (energy_hydrology.jl:1282-1327)
    assert np.allclose(a.total_water, ((Y.theta_l + Y.theta_i * E["rho_i"] / E["rho_l"]) * dz).sum(axis=1), rtol=1e-14)
    assert np.allclose(a.total_energy, (Y.rho_e_int * dz).sum(axis=1), rtol=1e-14)


def test_phase_change_conserves_water_mass():
    """energy_hydrology.jl:903-905: the two tendencies cancel in liquid-equivalent water,
    d(theta_l) + (rho_i / rho_l) d(theta_i) = 0, and the source ADDS into dY."""
    w, xp, P, Y = _problem(seed=5)
    Xp = P.explicit_params(**xp)
    a = P.new_aux()
    P.update_aux(Xp, Y, a)
    sl, si = np.zeros_like(Y.theta_l), np.zeros_like(Y.theta_l)
    P.phase_change(Xp, Y, a, sl, si)
    assert np.any(sl != 0.0)
    assert np.allclose(sl + E["rho_i"] / E["rho_l"] * si, 0.0, atol=4 * EPS * np.abs(sl).max())
    dl, di = np.full_like(Y.theta_l, 1.0), np.full_like(Y.theta_l, -2.0)
    P.phase_change(Xp, Y, a, dl, di)
    assert np.array_equal(dl, 1.0 + sl) and np.array_equal(di, -2.0 + si)
    # frozen cells with liquid above theta_star freeze (source > 0 -> d theta_l < 0)
    frozen = (a.T < a.Tf_depressed - 1.0) & (Y.theta_i == 0.0)
    if frozen.any():
        assert np.all(sl[frozen] <= 0.0)


# ----------------------------------------------------------------------------
# TOPMODEL runoff: test/standalone/Soil/runoff.jl
# ----------------------------------------------------------------------------
def _runoff_setup(model, ncol=7, N=15):
    """the reference test's configuration (runoff.jl:91-176): depth 50 m, theta_l = 0.6 - 0.3/50 (z + 50),
    nu = 0.5, theta_r = 0, K_sat = 1e-6, alpha = 0.2, n = 2.2, f_max = 0.5, f_over = 3.28, R_sb = 1.484e-7"""
    import climaland_b200  # noqa: F401
    from climaland_b200 import workloads
    z_f, z_c = workloads.stretched_grid(N, depth=50.0, dz_top=0.05)
    lat = np.linspace(-80.0, 80.0, ncol)
    precip = -1e-6 + 5e-7 * np.sin(lat / (90.0 * 2 * np.pi))
    fields = dict(nu=0.5, theta_r=0.0, K_sat=1e-6, S_s=1e-3, hcm_a=0.2, hcm_b=2.2, hcm_m=1 - 1 / 2.2)
    if model == orc.ENERGY_HYDROLOGY:
        fields["rho_c_ds"] = 2e6 * 0.5
    P = orc.Problem(model=model, z_f=z_f, z_c=z_c, ncol=ncol, **fields)
    Y = P.new_state()
    Y.theta_l[...] = 0.6 - 0.3 / 50.0 * (z_c + 50.0)
    return P, Y, z_f, z_c, precip


def test_topmodel_runoff_richards():
    """runoff.jl:177-214: h∇ is the column integral of heaviside(theta_l - nu) (theta_l - theta_r)/(nu - theta_r),
    R_ss = topmodel_ss_flux(R_sb, f_over, depth - h∇), the infiltration capacity is -K_sat, infiltration =
    topmodel_surface_infiltration(...), R_s = |precip - infiltration|"""
    P, Y, z_f, z_c, precip = _runoff_setup(orc.RICHARDS)
    f_max, f_over, R_sb, depth = 0.5, 3.28, 1.484e-4 / 1000, 50.0
    out = P.update_runoff(Y, precip, f_max, f_over, R_sb, depth)
    dz = np.diff(z_f)
    w = np.where(Y.theta_l - 0.5 > EPS, 1.0, 0.0) * (Y.theta_l - 0.0) / (0.5 - 0.0)
    h = (w * dz).sum(axis=1)
    assert np.allclose(out.h_grad, h, rtol=1e-15) and np.all(h > 0) and np.all(h < depth)
    assert np.array_equal(out.is_saturated, w)
    assert np.array_equal(out.R_ss, [L.orc_topmodel_ss_flux(R_sb, f_over, depth - x) for x in out.h_grad])
    assert np.allclose(out.R_ss, R_sb * np.exp(-f_over * (depth - out.h_grad)), rtol=1e-15)
    inf = [L.orc_topmodel_surface_infiltration(f_max, f_over, depth - x, -1e-6, pr) for x, pr in zip(out.h_grad, precip)]
    assert np.array_equal(out.infiltration, inf)
    f_sat = np.minimum(f_max * np.exp(-f_over / 2 * (depth - out.h_grad)), 1.0)
    assert np.allclose(out.infiltration, (1 - f_sat) * np.maximum(-1e-6, precip), rtol=1e-15)
    assert np.array_equal(out.R_s, np.abs(precip - out.infiltration))


def test_topmodel_runoff_energy_hydrology():
    """runoff.jl:216-300: with ice the surface-runoff saturation counts theta_l + theta_i, the subsurface one
    only theta_l; the infiltration capacity is -K_sat * impedance * viscosity at the top cell; R_ess is the
    saturated layers' liquid energy integral times R_ss / max(h∇, eps)."""
    P, Y, z_f, z_c, precip = _runoff_setup(orc.ENERGY_HYDROLOGY)
    ncol, N = Y.theta_l.shape
    Y.theta_i[...] = 0.06                            # ice pushes more levels above saturation, for R_s only
    xp = dict(kappa_dry=1.0, kappa_sat_unfrozen=1.5, kappa_sat_frozen=2.5, nu_ss_om=0.1, nu_ss_quartz=0.2, nu_ss_gravel=0.1)
    Xp = P.explicit_params(**xp)
    rc = 1e6 + np.minimum(0.5 - Y.theta_i, Y.theta_l) * E["rho_l"] * E["cp_l"] + Y.theta_i * E["rho_i"] * E["cp_i"]
    T = 270.0 + 10.0 * np.linspace(0, 1, N)[None, :] * np.ones((ncol, 1))
    Y.rho_e_int[...] = rc * (T - E["T_ref"]) - Y.theta_i * E["rho_i"] * E["LH_f0"]
    a = P.new_aux()
    P.update_aux(Xp, Y, a)
    f_max, f_over, R_sb, depth = 0.5, 3.28, 1.484e-4 / 1000, 50.0
    out = P.update_runoff(Y, precip, f_max, f_over, R_sb, depth, X=Xp, a=a)
    dz = np.diff(z_f)
    w_liq = np.where((Y.theta_l - 0.0) - 0.5 > EPS, 1.0, 0.0) * Y.theta_l / 0.5
    w_all = np.where((Y.theta_l + Y.theta_i) - 0.5 > EPS, 1.0, 0.0) * (Y.theta_l + Y.theta_i) / 0.5
    assert np.array_equal(out.is_saturated, w_liq)
    h_liq, h_all = (w_liq * dz).sum(axis=1), (w_all * dz).sum(axis=1)
    assert np.all(h_all > h_liq)
    assert np.allclose(out.h_grad, h_liq, rtol=1e-15)
    top = N - 1
    ic = np.array([-1e-6 * L.orc_impedance_factor(Y.theta_i[c, top] / (a.theta_l[c, top] + Y.theta_i[c, top] - 0.0), X["Omega"])
                   * L.orc_viscosity_factor(a.T[c, top], X["gamma"], X["gammaT_ref"]) for c in range(ncol)])
    f_sat = np.minimum(f_max * np.exp(-f_over / 2 * (depth - h_all)), 1.0)
    assert np.allclose(out.infiltration, (1 - f_sat) * np.maximum(ic, precip), rtol=1e-14)
    assert np.array_equal(out.R_s, np.abs(precip - out.infiltration))
    e_l = E["rho_l"] * E["cp_l"] * (a.T - E["T_ref"])
    assert np.allclose(out.R_ess, (w_liq * e_l * dz).sum(axis=1) * out.R_ss / np.maximum(h_liq, EPS), rtol=1e-13)
