"""The reference's own known-answer tests for the implicit soil path, run through the CUDA
library behind the host mirror of the reference interface (climaland.jl_b200/soil.py).  Each
test names the reference test it transcribes (paths relative to the ClimaLand.jl tree);
tolerances are the reference's (`≈` in Julia is rtol = sqrt(eps))."""
import math
from types import SimpleNamespace

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EPS = np.finfo(np.float64).eps
RTOL = math.sqrt(EPS)


def approx(a, b):
    return np.allclose(a, b, rtol=RTOL, atol=0.0)


def cl():
    import climaland_b200
    return climaland_b200


def oracle_lib():
    import oracle as orc
    return orc, orc.lib()


CLAY = dict(ν=0.495, K_sat=0.0443 / 3600 / 100, S_s=1e-3, θ_r=0.124)
CLAY_VG = dict(α=2.6, n=1.43)


def _clay_K_dpsi():
    orc, L = oracle_lib()
    m = 1 - 1 / CLAY_VG["n"]
    K = L.orc_vg_hydraulic_conductivity(m, CLAY["K_sat"], L.orc_effective_saturation(CLAY["ν"], 0.24, CLAY["θ_r"]))
    d = L.orc_vg_dpsidtheta(CLAY_VG["α"], CLAY_VG["n"], m, 0.24, CLAY["ν"], CLAY["θ_r"], CLAY["S_s"])
    return K, d


@pytest.mark.parametrize("ncol", [1, 3])  # Column and HybridBox in the reference
def test_richards_jacobian_moisture_bc(ncol):
    """test/shared_utilities/implicit_timestepping/richards_model.jl:16-141"""
    C = cl()
    params = C.RichardsParameters(hydrology_cm=C.vanGenuchten(**CLAY_VG), **CLAY)
    domain = C.Column(zlim=(-1.5, 0.0), nelements=150, ncol=ncol)
    bcs = dict(top=C.MoistureStateBC(lambda p, t: CLAY["ν"] - 1e-3), bottom=C.FreeDrainage())
    soil = C.RichardsModel(parameters=params, domain=domain, boundary_conditions=bcs, sources=())
    Y, p, coords = C.initialize(soil)
    Y.soil.ϑ_l[...] = 0.24
    uic = C.make_update_implicit_cache(soil)
    uic(p, Y, 0.0)
    jacobian = C.initialize_jacobian(soil)
    # solver type / algorithm / keys as asserted at richards_model.jl:70-82
    assert jacobian.solver_algorithm == "BlockDiagonalSolve"
    assert jacobian.keys == [("soil.ϑ_l", "soil.ϑ_l")]
    C.make_compute_jacobian(soil)(jacobian, Y, p, 1.0, 0.0)
    lo, di, up = jacobian.block(("soil.ϑ_l", "soil.ϑ_l"))
    K, d = _clay_K_dpsi()
    dz = 0.01
    for c in range(ncol):
        assert approx(lo[c, 1:], up[c, :-1])
        assert lo[c, 0] == 0.0 and up[c, -1] == 0.0
        assert approx(lo[c, 1:], 1.0 * (K / dz**2 * d))
        assert approx(di[c, 0], 1.0 * (-K / dz**2 * d) - 1)
        assert approx(di[c, 1:-1], 1.0 * (-2 * K / dz**2 * d) - 1)
        assert approx(di[c, -1], 1.0 * (-K / dz**2 * d - K / (dz * dz / 2) * d) - 1)


def test_richards_jacobian_flux_bc():
    """richards_model.jl:143-227: top WaterFluxBC(-K_sat), bottom FreeDrainage"""
    C = cl()
    params = C.RichardsParameters(hydrology_cm=C.vanGenuchten(**CLAY_VG), **CLAY)
    domain = C.Column(zlim=(-1.5, 0.0), nelements=150, ncol=2)
    bcs = dict(top=C.WaterFluxBC(lambda p, t: -CLAY["K_sat"]), bottom=C.FreeDrainage())
    soil = C.RichardsModel(parameters=params, domain=domain, boundary_conditions=bcs, sources=())
    Y, p, _ = C.initialize(soil)
    Y.soil.ϑ_l[...] = 0.24
    C.make_update_implicit_cache(soil)(p, Y, 0.0)
    jac = C.initialize_jacobian(soil)
    C.make_compute_jacobian(soil)(jac, Y, p, 1.0, 0.0)
    _, di, _ = jac.block(("soil.ϑ_l", "soil.ϑ_l"))
    K, d = _clay_K_dpsi()
    dz = 0.01
    for c in range(2):
        assert approx(di[c, 0], (-K / dz**2 * d) - 1)
        assert approx(di[c, 1:-1], (-2 * K / dz**2 * d) - 1)
        assert approx(di[c, -1], (-K / dz**2 * d) - 1)


def test_energy_hydrology_jacobian_flux_bc():
    """test/shared_utilities/implicit_timestepping/energy_hydrology_model.jl:16-173 (K, kappa are the
    lagged cache inputs of EnergyHydrology; entries are checked with the same formulas in K_ic, kappa_ic,
    the off-diagonal block WITH its -I, :163-172)"""
    C = cl()
    orc, L = oracle_lib()
    E = orc.EARTH
    Kvg, d = _clay_K_dpsi()
    K_ic = L.orc_impedance_factor(0.0, 7.0) * L.orc_viscosity_factor(280.0, 2.64e-2, 288.0) * Kvg
    kappa_ic = 1.37
    rho_c_ds = 2.3e6 * (1 - CLAY["ν"])
    params = C.EnergyHydrologyParameters(hydrology_cm=C.vanGenuchten(**CLAY_VG), ρc_ds=rho_c_ds, **CLAY)
    zero = C.WaterHeatBC(water=C.WaterFluxBC(0.0), heat=C.HeatFluxBC(0.0))
    soil = C.EnergyHydrology(parameters=params, domain=C.Column(zlim=(-1.5, 0.0), nelements=150, ncol=2),
                             boundary_conditions=dict(top=zero, bottom=zero), sources=())
    Y, p, _ = C.initialize(soil)
    Y.soil.ϑ_l[...] = 0.24
    Y.soil.θ_i[...] = 0.0
    rho_c_s = L.orc_volumetric_heat_capacity(0.24, 0.0, rho_c_ds, E["rho_l"], E["cp_l"], E["rho_i"], E["cp_i"])
    Y.soil.ρe_int[...] = L.orc_volumetric_internal_energy(0.0, rho_c_s, 280.0, E["rho_i"], E["T_ref"], E["LH_f0"])
    p.soil.K[...], p.soil.κ[...], p.soil.θ_l[...] = K_ic, kappa_ic, 0.24
    C.make_update_implicit_cache(soil)(p, Y, 0.0)
    assert approx(p.soil.T, 280.0)
    jac = C.initialize_jacobian(soil)
    assert jac.solver_algorithm == "BlockLowerTriangularSolve(soil.ϑ_l)"
    assert ("soil.ρe_int", "soil.ϑ_l") in jac.keys and ("soil.ϑ_l", "soil.ρe_int") not in jac.keys
    C.make_compute_jacobian(soil)(jac, Y, p, 1.0, 0.0)
    dz = 0.01
    dTdrho = 1 / rho_c_s
    e_liq = L.orc_volumetric_internal_energy_liq(280.0, E["rho_l"], E["cp_l"], E["T_ref"])
    for key, A, coef in ((("soil.ϑ_l", "soil.ϑ_l"), K_ic, d), (("soil.ρe_int", "soil.ρe_int"), kappa_ic, dTdrho),
                         (("soil.ρe_int", "soil.ϑ_l"), e_liq * K_ic, d)):
        _, di, _ = jac.block(key)
        for c in range(2):
            assert approx(di[c, 0], (-A / dz**2 * coef) - 1)
            assert approx(di[c, 1:-1], (-2 * A / dz**2 * coef) - 1)
            assert approx(di[c, -1], (-A / dz**2 * coef) - 1)


def test_richards_hydrostatic_zero_tendency():
    """test/standalone/Soil/soiltest.jl:15-90"""
    C = cl()
    N, zmin = 50, -10.0
    nu, n, a, theta_r = 0.495, 2.0, 2.6, 0.0
    params = C.RichardsParameters(hydrology_cm=C.vanGenuchten(α=a, n=n), ν=nu, K_sat=0.0443 / 3600 / 100, S_s=1e-3,
                                  θ_r=theta_r)
    domain = C.Column(zlim=(zmin, 0.0), nelements=N)
    soil = C.RichardsModel(parameters=params, domain=domain,
                           boundary_conditions=dict(top=C.WaterFluxBC(0.0), bottom=C.WaterFluxBC(0.0)))
    Y, p, coords = C.initialize(soil)
    z = coords.subsurface.z
    S = (1 + (a * (z - zmin)) ** n) ** (-(1 - 1 / n))
    Y.soil.ϑ_l[...] = S * (nu - theta_r) + theta_r
    C.make_update_implicit_cache(soil)(p, Y, 0.0)
    dY, _, _ = C.initialize(soil)
    C.make_compute_imp_tendency(soil)(dY, Y, p, 0.0)
    assert np.mean(dY.soil.ϑ_l) < EPS
    assert np.mean(p.soil.ψ + z + 10.0) < 2 * EPS
    assert np.max(np.abs(dY.soil.ϑ_l)) < 1e-14


def test_energy_hydrology_tendency_matches_hand_built_flux_formula():
    """soiltest.jl:97-406: implicit tendency against the hand-built face-flux formula, 1e2*eps"""
    C = cl()
    orc, L = oracle_lib()
    E = orc.EARTH
    N, zmin = 200, -1.0
    nu, n, a, theta_r, S_s = 0.495, 2.0, 2.6, 0.1, 1e-3
    m = 1 - 1 / n
    K_sat = 0.0443 / 3600 / 100
    domain = C.Column(zlim=(zmin, 0.0), nelements=N)
    z = domain.z_c
    dz = 1.0 / N
    theta = nu / 2 + nu / 4 * (z + 0.5) ** 2
    T = 280.0 + 0.5 * (z + 0.5) ** 2 * 10
    Kc = np.array([L.orc_vg_hydraulic_conductivity(m, K_sat, L.orc_effective_saturation(nu, t, theta_r)) for t in theta])
    kappa = 1.0 + 0.3 * np.sin(3 * z)
    rho_c_ds = 2e6 * (1 - nu)
    params = C.EnergyHydrologyParameters(hydrology_cm=C.vanGenuchten(α=a, n=n), ν=nu, K_sat=K_sat, S_s=S_s, θ_r=theta_r,
                                         ρc_ds=rho_c_ds)
    zero = C.WaterHeatBC(water=C.WaterFluxBC(0.0), heat=C.HeatFluxBC(0.0))
    soil = C.EnergyHydrology(parameters=params, domain=domain, boundary_conditions=dict(top=zero, bottom=zero))
    Y, p, _ = C.initialize(soil)
    Y.soil.ϑ_l[0] = theta
    rho_c_s = rho_c_ds + theta * E["rho_l"] * E["cp_l"]
    Y.soil.ρe_int[0] = rho_c_s * (T - E["T_ref"])
    p.soil.K[0], p.soil.κ[0], p.soil.θ_l[0] = Kc, kappa, theta
    C.make_update_implicit_cache(soil)(p, Y, 0.0)
    dY, _, _ = C.initialize(soil)
    C.make_compute_imp_tendency(soil)(dY, Y, p, 0.0)
    psi = np.array([L.orc_vg_pressure_head(a, n, m, theta_r, t, nu, S_s) for t in theta])
    h = psi + z
    flux = np.concatenate([[0.0], -0.5 * (Kc[1:] + Kc[:-1]) * (h[1:] - h[:-1]) / dz, [0.0]])
    expected = -(flux[1:] - flux[:-1]) / dz
    assert np.mean(np.abs(expected - dY.soil.ϑ_l[0])) / nu < 1e2 * EPS
    Tc = p.soil.T[0]
    e_l = E["rho_l"] * E["cp_l"] * (Tc - E["T_ref"])
    eK_face = 0.5 * ((e_l * Kc)[1:] + (e_l * Kc)[:-1])
    k_face = 0.5 * (kappa[1:] + kappa[:-1])
    flux = np.concatenate([[0.0], -k_face * (Tc[1:] - Tc[:-1]) / dz - eK_face * (h[1:] - h[:-1]) / dz, [0.0]])
    expected = -(flux[1:] - flux[:-1]) / dz
    assert np.mean(np.abs(expected - dY.soil.ρe_int[0])) / np.median(Y.soil.ρe_int[0]) < 1e2 * EPS
    assert np.all(dY.soil.θ_i == 0.0)


@pytest.mark.parametrize("model", ["richards", "energy_hydrology"])
def test_flux_bc_conservation_signs(model):
    """test/standalone/Soil/conservation.jl:103-151, 218-258: d(∫F)/dt = -(F_top - F_bot) = -2 and the
    column-integrated tendency equals it."""
    C = cl()
    N = 20
    cm = C.vanGenuchten(α=2.6, n=2.0)
    common = dict(ν=0.495, K_sat=0.0443 / 3600 / 100, S_s=1e-3, θ_r=0.0)
    domain = C.Column(zlim=(-1.0, 0.0), nelements=N, ncol=2)
    if model == "richards":
        soil = C.RichardsModel(parameters=C.RichardsParameters(hydrology_cm=cm, **common), domain=domain,
                               boundary_conditions=dict(top=C.WaterFluxBC(1.0), bottom=C.WaterFluxBC(-1.0)))
    else:
        bc = lambda v: C.WaterHeatBC(water=C.WaterFluxBC(v), heat=C.HeatFluxBC(v))  # noqa: E731
        soil = C.EnergyHydrology(parameters=C.EnergyHydrologyParameters(hydrology_cm=cm, ρc_ds=1e6, **common),
                                 domain=domain, boundary_conditions=dict(top=bc(1.0), bottom=bc(-1.0)))
    Y, p, _ = C.initialize(soil)
    Y.soil.ϑ_l[...] = 0.495 / 2
    if model == "energy_hydrology":
        Y.soil.ρe_int[...] = 2.0e7
        p.soil.K[...], p.soil.κ[...], p.soil.θ_l[...] = 1e-7, 1.5, 0.2475
    C.make_update_implicit_cache(soil)(p, Y, 0.0)
    dY, _, _ = C.initialize(soil)
    C.make_compute_imp_tendency(soil)(dY, Y, p, 0.0)
    dz = np.diff(domain.z_f)
    assert np.all(dY.soil.ᶠF_vol_liq_water_dt == -2.0)
    assert np.allclose(dY.soil.ϑ_l @ dz, -2.0, rtol=1e-12)
    if model == "energy_hydrology":
        assert np.all(dY.soil.ᶠF_e_dt == -2.0)
        assert np.allclose(dY.soil.ρe_int @ dz, -2.0, rtol=1e-9)


def test_moisture_state_bc_flux_and_free_drainage():
    """test/standalone/Soil/soil_bc.jl:98-133 (state -> flux, diffusive_flux(K_c, psi_bc + dz, psi_c, dz)) and
    boundary_conditions.jl:340-353 (bottom_bc = -K_1), conservation.jl:127-133 (total water)"""
    C = cl()
    orc, L = oracle_lib()
    N = 50
    nu, n, a, m = 0.495, 2.0, 2.6, 0.5
    K_sat, S_s = 0.0443 / 3600 / 100, 1e-3
    params = C.RichardsParameters(hydrology_cm=C.vanGenuchten(α=a, n=n), ν=nu, K_sat=K_sat, S_s=S_s, θ_r=0.0)
    domain = C.Column(zlim=(-10.0, 0.0), nelements=N)
    soil = C.RichardsModel(parameters=params, domain=domain,
                           boundary_conditions=dict(top=C.MoistureStateBC(nu / 2), bottom=C.MoistureStateBC(nu / 2)))
    Y, p, _ = C.initialize(soil)
    Y.soil.ϑ_l[...] = nu / 3
    C.make_update_implicit_cache(soil)(p, Y, 0.0)
    dz = 10.0 / N / 2.0
    K_c = L.orc_vg_hydraulic_conductivity(m, K_sat, L.orc_effective_saturation(nu, nu / 3, 0.0))
    psi_bc = L.orc_vg_pressure_head(a, n, m, 0.0, nu / 2, nu, S_s)
    psi_c = L.orc_vg_pressure_head(a, n, m, 0.0, nu / 3, nu, S_s)
    assert approx(p.soil.top_bc[0], -K_c * ((psi_bc - psi_c + dz) / dz))
    assert approx(p.soil.bottom_bc[0], -K_c * ((psi_c + dz - psi_bc) / dz))
    assert approx(p.soil.dfluxBCdY[0], K_c * L.orc_vg_dpsidtheta(a, n, m, nu / 3, nu, 0.0, S_s) / dz)
    assert approx(p.soil.total_water[0], nu / 3 * 10.0)
    # free drainage
    soil2 = C.RichardsModel(parameters=params, domain=domain,
                            boundary_conditions=dict(top=C.MoistureStateBC(0.4), bottom=C.FreeDrainage()))
    Y2, p2, _ = C.initialize(soil2)
    Y2.soil.ϑ_l[...] = nu / 2
    C.make_update_implicit_cache(soil2)(p2, Y2, 0.0)
    assert np.all(p2.soil.bottom_bc == -p2.soil.K[:, 0])


def test_mask_leaves_inactive_columns_untouched():
    """test/standalone/Soil/mask_test.jl:53-61, test/integrated/full_land.jl:586-637: tendency, Jacobian
    solve and state of masked (ocean) columns are never written."""
    C = cl()
    ncol, N = 40, 15
    active = np.array([c for c in range(ncol) if c % 3 != 1])
    params = C.RichardsParameters(hydrology_cm=C.vanGenuchten(**CLAY_VG), **CLAY)
    domain = C.Column(zlim=(-1.5, 0.0), nelements=N, ncol=ncol, active_columns=active)
    soil = C.RichardsModel(parameters=params, domain=domain,
                           boundary_conditions=dict(top=C.WaterFluxBC(-1e-7), bottom=C.FreeDrainage()))
    Y, p, _ = C.initialize(soil)
    rng = np.random.default_rng(0)
    Y.soil.ϑ_l[...] = rng.uniform(0.2, 0.45, Y.soil.ϑ_l.shape)
    C.make_update_implicit_cache(soil)(p, Y, 0.0)
    dY, _, _ = C.initialize(soil)
    SENTINEL = -12345.0
    dY.soil.ϑ_l[...] = SENTINEL
    C.make_compute_imp_tendency(soil)(dY, Y, p, 0.0)
    inactive = np.setdiff1d(np.arange(ncol), active)
    assert np.all(dY.soil.ϑ_l[inactive] == SENTINEL)
    assert np.all(dY.soil.ϑ_l[active] != SENTINEL)
    # the fused stage through LandSimulation: inactive columns keep their state bit for bit
    sim = C.LandSimulation(0.0, 1800.0, 1800.0, soil,
                           timestepper=C.IMEXAlgorithm("ARS111", C.FusedSoilNewton(max_iters=2)))
    sim.Y.soil.ϑ_l[...] = Y.soil.ϑ_l
    before = sim.Y.soil.ϑ_l.copy()
    sim.step()
    assert np.array_equal(sim.Y.soil.ϑ_l[inactive], before[inactive])
    assert np.all(sim.Y.soil.ϑ_l[active] != before[active])
    assert sim.stats["nan_count"] == 0
