"""GPU parity of the explicit-stage soil kernels (update_aux! of EnergyHydrology and the PhaseChange
source; energy_hydrology.jl:722-906) against the CPU oracle on identical seeded inputs,
through the C ABI.  Tolerance: 1e-12 norm-wise relative per call (BASELINE.json north_star)."""
import numpy as np
import pytest

from helpers import assert_close, cuda_solver, oracle_problem

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _setup(ncol, N, seed, closure, math_mode, layout, explicit_kernel=0, ice=True):
    import climaland_b200 as cl
    from climaland_b200 import workloads
    w = workloads.make_workload("energy_hydrology", ncol, N=N, seed=seed, topmodel=False, ice=ice)
    if closure == 1:
        from test_cuda_hooks_parity import _to_brooks_corey
        w = _to_brooks_corey(w)
    xp = workloads.make_explicit_params(w, seed)
    P, Y, cache = oracle_problem(w, closure=closure, nthreads=4)
    Xp = P.explicit_params(**xp)
    a = P.new_aux()
    P.update_aux(Xp, Y, a)
    s = cuda_solver(w, closure=closure, math_mode=math_mode, layout=layout)
    s.set_option("explicit_kernel", explicit_kernel)
    for k, v in xp.items():
        s.set(k, v)
    s.set_explicit_params(**workloads.EXPLICIT_SCALARS)
    P.cache = cache  # boundary fluxes of the workload, for the implicit stage
    return cl, w, P, Xp, Y, a, s


AUX_FIELDS = [("theta_l_lag", "theta_l"), ("kappa_lag", "kappa"), ("k_lag", "K"), ("p_t", "T"), ("p_psi", "psi"),
              ("p_tf_depressed", "Tf_depressed"), ("total_water", "total_water"), ("total_energy", "total_energy")]


@pytest.mark.parametrize("layout", [1, 2], ids=["colfast", "levfast"])
@pytest.mark.parametrize("math_mode", [0, 1], ids=["fast", "libm"])
@pytest.mark.parametrize("closure,N,ncol", [(0, 15, 1000), (0, 50, 130), (1, 15, 257)])
def test_update_aux(closure, N, ncol, math_mode, layout):
    cl, w, P, Xp, Y, a, s = _setup(ncol, N, 11, closure, math_mode, layout)
    s.update_aux()
    for dev, orc_name in AUX_FIELDS:
        assert_close(s.get(dev), getattr(a, orc_name), TOL, orc_name)
    s.close()


@pytest.mark.parametrize("fused", [False, True], ids=["separate", "fused"])
@pytest.mark.parametrize("math_mode", [0, 1], ids=["fast", "libm"])
@pytest.mark.parametrize("closure,N,ncol", [(0, 15, 1000), (1, 15, 257)])
def test_phase_change_source(closure, N, ncol, math_mode, fused):
    cl, w, P, Xp, Y, a, s = _setup(ncol, N, 12, closure, math_mode, 0)
    rng = np.random.default_rng(0)
    d0l, d0i = rng.normal(0, 1e-8, Y.theta_l.shape), rng.normal(0, 1e-8, Y.theta_l.shape)
    dl, di = d0l.copy(), d0i.copy()
    P.phase_change(Xp, Y, a, dl, di)
    s.set("dye_theta_l", d0l)
    s.set("dye_theta_i", d0i)
    if fused:
        s.update_aux_and_phase_change()
    else:
        s.update_aux()
        s.phase_change_source()
    # the source itself (what was added) to 1e-12 of its own scale, and the accumulated tendency
    assert_close(s.get("dye_theta_l") - d0l, dl - d0l, 1e-9, "source (difference of accumulations)")
    # element-wise 2e-12: the source is (theta_l - theta_star)/tau with theta_star a power of the freezing-point
    # depression (two roundings of <= 2 ulp in FAST mode on top of the difference); measured 1.1e-12
    assert_close(s.get("dye_theta_l"), dl, TOL, "dY.theta_l", elem_tol=2e-12)
    assert_close(s.get("dye_theta_i"), di, TOL, "dY.theta_i", elem_tol=2e-12)
    s.close()


@pytest.mark.parametrize("ice", [True, False], ids=["ice", "noice"])
@pytest.mark.parametrize("layout", [1, 2], ids=["colfast", "levfast"])
@pytest.mark.parametrize("closure,N,ncol", [(0, 15, 1003), (1, 15, 257), (0, 50, 131)])
def test_percell_and_warp_uniform_kernels_agree(closure, N, ncol, layout, ice):
    """CLB_OPT_EXPLICIT_KERNEL: the kernel that follows the reference's case distinctions cell by cell (1) and the one
    with warp-uniform control flow (0, the default; what every FAST case above ran) evaluate the same formulas: equal
    up to the last bits of one exp(log k_u), both within the per-call bar of the oracle; with ice in some cells of a
    warp and with none anywhere (the shortcut of the whole warp); column counts that leave a partial last warp."""
    out = {}
    for kern in (0, 1):
        cl, w, P, Xp, Y, a, s = _setup(ncol, N, 17, closure, 0, layout, explicit_kernel=kern, ice=ice)
        dl, di = np.zeros_like(Y.theta_l), np.zeros_like(Y.theta_l)
        P.phase_change(Xp, Y, a, dl, di)
        s.update_aux_and_phase_change()
        out[kern] = {dev: s.get(dev) for dev, _ in AUX_FIELDS}
        out[kern].update(dye_theta_l=s.get("dye_theta_l"), dye_theta_i=s.get("dye_theta_i"))
        for dev, orc_name in AUX_FIELDS:
            assert_close(out[kern][dev], getattr(a, orc_name), TOL, orc_name)
        assert_close(out[kern]["dye_theta_l"], dl, TOL, "source theta_l", elem_tol=2e-12)
        s.close()
    for k in out[0]:  # the source: a difference, 1 / alpha and g / LH as multiplications by reciprocals in kernel 0
        assert_close(out[0][k], out[1][k], 1e-13, "source " + k if k.startswith("dye") else k, elem_tol=2e-12)


def test_phase_change_alone_matches_source_scale():
    """zero initial tendency: the kernel's source against the oracle's at 1e-12 of the source's own scale"""
    cl, w, P, Xp, Y, a, s = _setup(2000, 15, 13, 0, 0, 0)
    dl, di = np.zeros_like(Y.theta_l), np.zeros_like(Y.theta_l)
    P.phase_change(Xp, Y, a, dl, di)
    s.update_aux_and_phase_change()
    assert np.abs(dl).max() > 0.0
    assert_close(s.get("dye_theta_l"), dl, TOL, "source theta_l")
    assert_close(s.get("dye_theta_i"), di, TOL, "source theta_i")
    s.close()


def test_explicit_stage_feeds_the_implicit_stage():
    """update_aux! writes the lagged inputs of the implicit stage in place: a fused implicit stage after
    it matches the oracle's stage run on the oracle's own aux values."""
    cl, w, P, Xp, Y, a, s = _setup(512, 15, 14, 0, 0, 0)
    s.update_aux()
    P.set("K_lag", a.K)
    P.set("kappa_lag", a.kappa)
    P.set("theta_l_lag", a.theta_l)
    U = Y.copy()
    P.implicit_step(U, 900.0, 3, p=P.cache)
    s.implicit_step(900.0, 3)
    assert_close(s.get("y_theta_l"), U.theta_l, TOL, "theta_l after the stage")
    assert_close(s.get("y_rho_e_int"), U.rho_e_int, TOL, "rho_e_int after the stage")
    s.close()


def test_errors():
    import climaland_b200 as cl
    from climaland_b200 import workloads
    w = workloads.make_workload("energy_hydrology", 64, N=15, seed=1)
    s = cuda_solver(w)
    with pytest.raises(cl.ClbError, match="clb_set_explicit_params"):
        s.update_aux()
    s.set_explicit_params(**workloads.EXPLICIT_SCALARS)
    with pytest.raises(cl.ClbError, match="never set"):
        s.update_aux()
    s.close()
    wr = workloads.make_workload("richards", 64, N=15, seed=1)
    r = cuda_solver(wr)
    r.set_explicit_params(**workloads.EXPLICIT_SCALARS)
    with pytest.raises(cl.ClbError, match="EnergyHydrology only"):
        r.update_aux()
    r.close()


def test_mirror_update_aux_and_phase_change_read_like_the_reference():
    """test/standalone/Soil/soil_parameterizations.jl:283-430 ("Freezing and Thawing") through the
    host-side mirror of the reference interface: make_update_aux(model)(p, Y, t), then
    source!(dY, PhaseChange(), Y, p, model).  Columns hold theta_l = (0.11, 0.15, nu); T = 270 K
    freezes (source > 0, d theta_l < 0), T = 274 K without ice leaves the state alone."""
    import oracle as orc
    import climaland_b200 as C
    L, E, X = orc.lib(), orc.EARTH, orc.EXPLICIT_SCALARS
    nu, theta_r, a, n, S_s, K_sat = 0.2, 0.1, 2.0, 1.4, 1e-3, 1e-5
    m = 1.0 - 1.0 / n
    rho_c_ds, k_dry, k_u, k_f = 1.8e6, 1.1, 1.4, 2.3
    params = C.EnergyHydrologyParameters(hydrology_cm=C.vanGenuchten(α=a, n=n), ν=nu, K_sat=K_sat, S_s=S_s, θ_r=theta_r,
                                         ρc_ds=rho_c_ds, κ_dry=k_dry, κ_sat_unfrozen=k_u, κ_sat_frozen=k_f,
                                         ν_ss_om=0.1, ν_ss_quartz=0.1, ν_ss_gravel=0.1)
    zero = C.WaterHeatBC(water=C.WaterFluxBC(0.0), heat=C.HeatFluxBC(0.0))
    soil = C.EnergyHydrology(parameters=params, domain=C.Column(zlim=(-3.0, 0.0), nelements=3, ncol=3),
                             boundary_conditions=dict(top=zero, bottom=zero), sources=(C.PhaseChange(),))
    for T, sign in ((270.0, 1), (274.0, 0)):
        Y, p, _ = C.initialize(soil)
        Y.soil.ϑ_l[...] = np.array([0.11, 0.15, nu])[:, None]
        Y.soil.θ_i[...] = 0.0
        rc = np.array([L.orc_volumetric_heat_capacity(t, 0.0, rho_c_ds, E["rho_l"], E["cp_l"], E["rho_i"], E["cp_i"])
                       for t in (0.11, 0.15, nu)])[:, None]
        Y.soil.ρe_int[...] = rc * (T - E["T_ref"])
        C.make_update_aux(soil)(p, Y, 0.0)
        assert np.allclose(p.soil.T, T, rtol=1e-14) and np.allclose(p.soil.θ_l, Y.soil.ϑ_l)
        dY, _, _ = C.initialize(soil)
        C.make_phase_change_source(soil)(dY, C.PhaseChange(), Y, p, soil)
        for c, theta_l in enumerate((0.11, 0.15, nu)):
            tau = L.orc_thermal_time(rc[c, 0], 1.0, p.soil.κ[c, 0])
            want = L.orc_phase_change_source(0, a, n, m, theta_l, 0.0, p.soil.T[c, 0], tau, nu, theta_r, E["rho_i"], E["rho_l"],
                                             E["LH_f0"], X["T_freeze"], X["grav"])
            assert np.allclose(dY.soil.ϑ_l[c], -want, rtol=1e-12, atol=1e-12 * theta_l / tau)
            assert np.allclose(dY.soil.θ_i[c], E["rho_l"] / E["rho_i"] * want, rtol=1e-12, atol=1e-12 * theta_l / tau)
            if sign:
                assert want > 0.0 and np.all(dY.soil.ϑ_l[c] < 0.0)
            else:
                assert abs(want) <= 1.5e-8 * theta_l / tau
    soil.solver.close()


@pytest.mark.parametrize("math_mode", [0, 1], ids=["fast", "libm"])
@pytest.mark.parametrize("model,layout", [("richards", 1), ("energy_hydrology", 1), ("energy_hydrology", 2)])
def test_topmodel_runoff(model, layout, math_mode):
    """update_infiltration_water_flux!(p, ::TOPMODELRunoff, ...) (Runoff/Runoff.jl:234-283) against the oracle;
    its outputs are the lagged inputs of the implicit TOPMODEL source, left in place for the fused stage."""
    import climaland_b200  # noqa: F401
    from climaland_b200 import workloads
    ncol, N, depth = 1500, 15, 50.0
    w = workloads.make_workload(model, ncol, N=N, seed=21, topmodel=True)
    rng = np.random.default_rng(2)
    # saturate the lower part of a third of the columns so that h∇ > 0 there
    sat = rng.random(ncol) < 0.35
    w["y_theta_l"][sat, :5] = w["nu"][sat, :5] + 2e-3
    precip = -rng.uniform(0.0, 2e-6, ncol)
    f_max = rng.uniform(0.2, 0.6, ncol)
    f_over, R_sb = 3.28, 1.484e-4 / 1000
    P, Y, _ = oracle_problem(w, nthreads=4)
    s = cuda_solver(w, math_mode=math_mode, layout=layout)
    X = a = None
    if model == "energy_hydrology":
        xp = workloads.make_explicit_params(w, 21)
        X = P.explicit_params(**xp)
        a = P.new_aux()
        P.update_aux(X, Y, a)
        for k, v in xp.items():
            s.set(k, v)
        s.set_explicit_params(**workloads.EXPLICIT_SCALARS)
        s.update_aux()
    out = P.update_runoff(Y, precip, f_max, f_over, R_sb, depth, X=X, a=a)
    assert out.h_grad.max() > 0.0
    s.set("f_max", f_max)
    s.set("precip", precip)
    s.set_runoff_params(f_over=f_over, R_sb=R_sb, depth=depth)
    s.update_runoff()
    assert_close(s.get("is_saturated"), out.is_saturated, TOL, "is_saturated")
    for dev, name in (("h_grad", "h_grad"), ("r_ss", "R_ss"), ("infiltration", "infiltration"), ("r_s", "R_s")):
        assert_close(s.get(dev), getattr(out, name), TOL, name)
    if model == "energy_hydrology":
        assert_close(s.get("r_ess"), out.R_ess, TOL, "R_ess")
    s.close()


def test_explicit_rows_respect_the_land_sea_mask():
    """mask_test.jl:53-61 for the explicit-stage and SoilCO2 entry points: a handle over the active columns of a
    larger domain reads / writes only those columns of the caller's arrays and gives the compacted problem's
    values; inactive columns of an output array keep what they held."""
    import oracle as orc
    import climaland_b200 as cl
    from climaland_b200 import workloads
    ncol_total, N = 400, 15
    rng = np.random.default_rng(8)
    active = np.sort(rng.choice(ncol_total, 150, replace=False)).astype(np.int64)
    w = workloads.make_workload("energy_hydrology", ncol_total, N=N, seed=17, topmodel=False)
    xp = workloads.make_explicit_params(w, 17)
    s = cl.SoilColumnSolver(model=cl.ENERGY_HYDROLOGY, n_columns=active.size, n_columns_total=ncol_total, z_f=w["z_f"],
                            z_c=w["z_c"], active_columns=active)
    for k, v in {**w, **xp}.items():
        if k.lower() in cl.FIELDS:
            s.set(k, v)
    s.set_explicit_params(**workloads.EXPLICIT_SCALARS)
    s.update_aux()
    # the compacted problem on the oracle
    sub = {k: (v[active] if isinstance(v, np.ndarray) and v.shape[:1] == (ncol_total,) else v) for k, v in w.items()}
    sub["ncol"] = active.size
    P, Y, _ = oracle_problem(sub, nthreads=2)
    X = P.explicit_params(**{k: v[active] for k, v in xp.items()})
    a = P.new_aux()
    P.update_aux(X, Y, a)
    sentinel = -777.0
    for dev, name in (("kappa_lag", "kappa"), ("k_lag", "K"), ("p_tf_depressed", "Tf_depressed")):
        out = np.full((ncol_total, N), sentinel)
        s.get(dev, out)
        assert_close(out[active], getattr(a, name), TOL, name)
        inactive = np.setdiff1d(np.arange(ncol_total), active)
        assert np.all(out[inactive] == sentinel)
    tw = np.full(ncol_total, sentinel)
    s.get("total_water", tw)
    assert_close(tw[active], a.total_water, TOL, "total_water")
    assert np.all(np.delete(tw, active) == sentinel)
    # SoilCO2 stage on the same masked handle
    D, th = rng.uniform(1e-8, 2e-6, (ncol_total, N)), rng.uniform(0.02, 0.45, (ncol_total, N))
    C0 = rng.uniform(5e-5, 2e-3, (ncol_total, N))
    for name in ("co2", "o2"):
        s.set(f"{name}_y", C0)
        s.set(f"{name}_d", D)
        s.set(f"{name}_theta_eff", th)
    s.soilco2_implicit_step(1800.0, 3)
    S = P.co2_species(D[active], th[active])
    C = np.ascontiguousarray(C0[active])
    P.co2_implicit_step(S, C, np.zeros(active.size), np.zeros(active.size), 1800.0, 3)
    out = np.full((ncol_total, N), sentinel)
    s.get("co2_y", out)
    assert_close(out[active], C, TOL, "CO2 after the stage")
    assert np.all(np.delete(out, active, axis=0) == sentinel)
    s.close()


# ---- the atmosphere-driven top boundary, SurfaceRunoff, EnergyWaterFreeDrainage (second half) ----
def _atmos_forcing(ncol, seed):
    rng = np.random.default_rng(seed)
    return dict(precip=-rng.uniform(0, 2e-6, ncol), vapor_flux_liq=rng.normal(0, 1e-8, ncol), lhf=rng.normal(0, 50, ncol),
                shf=rng.normal(0, 30, ncol), r_n=rng.normal(-100, 50, ncol), t_air=rng.uniform(260, 300, ncol))


@pytest.mark.parametrize("math_mode", [0, 1], ids=["fast", "libm"])
@pytest.mark.parametrize("runoff_model", [0, 1, 2], ids=["norunoff", "surface", "topmodel"])
def test_atmos_driven_top_fluxes(runoff_model, math_mode):
    """clb_update_atmos_driven_fluxes against the oracle's soil_boundary_fluxes!(::AtmosDrivenFluxBC, ...)
    (boundary_conditions.jl:901-936, Runoff.jl:69-71, 129-148, 234-283)"""
    ncol = 777
    cl, w, P, Xp, Y, a, s = _setup(ncol, 15, 21, 0, math_mode, 0)
    F = _atmos_forcing(ncol, 5)
    sat_cols = np.arange(0, ncol, 5)  # a saturated surface in a fifth of the columns: everything runs off there
    Y.theta_l[sat_cols, -1] = (w["nu"] - Y.theta_i)[sat_cols, -1] + 1e-3
    P.update_aux(Xp, Y, a)
    s.set("y_theta_l", Y.theta_l)
    for k, v in F.items():
        s.set(k, v)
    f_max = np.random.default_rng(1).uniform(0.2, 0.6, ncol)
    s.set("f_max", f_max)
    s.set_runoff_params(f_over=3.28, R_sb=1.484e-7, depth=50.0)
    s.update_aux()
    s.update_atmos_driven_fluxes(runoff_model)
    if runoff_model == 2:
        R = P.update_runoff(Y, F["precip"], f_max, 3.28, 1.484e-7, 50.0, X=Xp, a=a)
        inf, sat, R_s = R.infiltration, R.is_saturated, R.R_s
    else:
        sat, inf, R_s = P.surface_runoff(Y, runoff_model, F["precip"], X=Xp, a=a)
    tw, th = P.atmos_driven_top_fluxes(inf, F["vapor_flux_liq"], F["lhf"], F["shf"], F["r_n"], F["t_air"])
    assert_close(s.get("infiltration"), inf, TOL, "infiltration")
    assert_close(s.get("top_bc_w"), tw, TOL, "top_bc.water")
    assert_close(s.get("top_bc_h"), th, TOL, "top_bc.heat")
    if runoff_model > 0:
        assert_close(s.get("is_saturated"), sat, TOL, "is_saturated")
        assert_close(s.get("r_s"), R_s, TOL, "R_s")
    if runoff_model == 1:
        assert np.all(s.get("infiltration")[sat_cols] == 0.0)
    s.close()


@pytest.mark.parametrize("math_mode", [0, 1], ids=["fast", "libm"])
def test_energy_water_free_drainage(math_mode):
    cl, w, P, Xp, Y, a, s = _setup(500, 15, 22, 0, math_mode, 0)
    s.update_aux()
    s.update_energy_water_free_drainage()
    bw, bh = P.energy_water_free_drainage(a)
    assert_close(s.get("bot_bc_w"), bw, TOL, "bottom_bc.water")
    assert_close(s.get("bot_bc_h"), bh, TOL, "bottom_bc.heat")
    s.close()
