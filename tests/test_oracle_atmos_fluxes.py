"""Pins the CPU oracle's atmosphere-driven top-flux assembly, SurfaceRunoff and EnergyWaterFreeDrainage (SURVEY 8f rank
2, second half) on the tests the reference holds for them.  The reference's tests are identities -- the computed
cache field must EQUAL the formula evaluated by the test -- so they are transcribed as such, on the reference's own
configurations where they are stated.  No GPU."""
import numpy as np

import oracle as orc
from helpers import oracle_problem

E = orc.EARTH


def _eh(ncol=64, seed=3):
    import climaland_b200  # noqa: F401
    from climaland_b200 import workloads
    w = workloads.make_workload("energy_hydrology", ncol, N=15, seed=seed, topmodel=False)
    xp = workloads.make_explicit_params(w, seed)
    P, Y, p = oracle_problem(w)
    X = P.explicit_params(**xp)
    a = P.new_aux()
    P.update_aux(X, Y, a)
    return w, P, Y, X, a


def test_atmos_driven_flux_identities():
    """test/standalone/Soil/climate_drivers.jl:240-262: with NoRunoff the computed top_bc.water == precip +
    vapor_flux_liq and top_bc.heat == R_n + lhf + shf + precip * volumetric_internal_energy_liq(T_atmos)"""
    w, P, Y, X, a = _eh()
    rng = np.random.default_rng(0)
    n = w["ncol"]
    precip = -rng.uniform(0, 1e-6, n)
    vap, lhf, shf, R_n = rng.normal(0, 1e-8, n), rng.normal(0, 50, n), rng.normal(0, 30, n), rng.normal(-100, 50, n)
    T_air = rng.uniform(260, 300, n)
    _, inf, _ = P.surface_runoff(Y, 0, precip)
    assert np.array_equal(inf, precip)  # Runoff.jl:69-71
    tw, th = P.atmos_driven_top_fluxes(inf, vap, lhf, shf, R_n, T_air)
    e_liq = np.array([orc.lib().orc_volumetric_internal_energy_liq(t, E["rho_l"], E["cp_l"], E["T_ref"]) for t in T_air])
    assert np.array_equal(tw, precip + vap)
    assert np.array_equal(th, R_n + lhf + shf + precip * e_liq)


def test_surface_runoff_richards_site_level():
    """test/standalone/Soil/runoff.jl:307-385 ("Richards model, Site level runoff"): nu = 0.5, K_sat = 1e-6, theta_l =
    0.6 - 0.3/50 (z + 50) (saturated towards the bottom, 0.3 at the surface); ic == -K_sat; infiltration ==
    surface_infiltration(ic, precip, is_saturated at the top centre); R_s == |precip - infiltration|"""
    import climaland_b200  # noqa: F401
    from climaland_b200 import workloads
    ncol = 40
    w = workloads.make_workload("richards", ncol, N=15, seed=1, depth=50.0)
    w["nu"][...] = 0.5
    w["theta_r"][...] = 0.0
    w["K_sat"][...] = 1e-6
    w["y_theta_l"][...] = 0.6 - 0.3 / 50.0 * (w["z_c"][None, :] + 50.0)
    # the reference's field is unsaturated at the surface everywhere; make a few columns saturated there as well
    w["y_theta_l"][::7, -1] = 0.51
    P, Y, p = oracle_problem(w)
    lat = np.linspace(-90, 90, ncol)
    precip = -1e-6 + 5e-7 * np.sin(lat / (90.0 * 2 * np.pi))
    sat, inf, R_s = P.surface_runoff(Y, 1, precip)
    assert np.array_equal(sat, (w["y_theta_l"] - w["nu"] > np.finfo(np.float64).eps).astype(float))  # heaviside, :432-434
    ic = -1e-6
    assert np.array_equal(inf, (1 - sat[:, -1]) * np.maximum(ic, precip))
    assert np.array_equal(R_s, np.abs(precip - inf))
    assert np.all(inf[::7] == 0.0) and np.all(R_s[::7] == np.abs(precip[::7]))  # saturated surface: all runs off


def test_surface_runoff_energy_hydrology_capacity():
    """soil_infiltration_capacity(model::EnergyHydrology, ...) (Runoff.jl:396-410): -K_sat x impedance x viscosity at the
    top cell, from p.soil.theta_l and p.soil.T"""
    w, P, Y, X, a = _eh(seed=4)
    precip = np.full(w["ncol"], -1.0)  # larger than any capacity: infiltration = capacity where the top is unsaturated
    sat, inf, R_s = P.surface_runoff(Y, 1, precip, X=X, a=a)
    L = orc.lib()
    S = orc.EXPLICIT_SCALARS
    top = -1
    want = np.array([-w["K_sat"][c, top] *
                     L.orc_impedance_factor(w["y_theta_i"][c, top] / (a.theta_l[c, top] + w["y_theta_i"][c, top] - w["theta_r"][c, top]), S["Omega"]) *
                     L.orc_viscosity_factor(a.T[c, top], S["gamma"], S["gammaT_ref"]) for c in range(w["ncol"])])
    assert np.array_equal(inf, (1 - sat[:, top]) * want)


def test_energy_water_free_drainage():
    """test/standalone/Soil/soil_bc.jl:218-266: bottom_bc.water == -K at level 1, bottom_bc.heat == that times
    volumetric_internal_energy_liq(T at level 1)"""
    w, P, Y, X, a = _eh(seed=5)
    bw, bh = P.energy_water_free_drainage(a)
    e1 = np.array([orc.lib().orc_volumetric_internal_energy_liq(t, E["rho_l"], E["cp_l"], E["T_ref"]) for t in a.T[:, 0]])
    assert np.array_equal(bw, -1 * a.K[:, 0])
    assert np.allclose(bh, bw * e1, rtol=1.5e-8, atol=0.0)  # the reference's own comparison is `≈`
