"""clb_soil_step_host: a WHOLE EnergyHydrology soil step (update_aux! + PhaseChange, TOPMODEL runoff, the explicit
update, the fused implicit stage) from and to host arrays holding the state at t_n.  Only the state and the forcing
cross PCIe; the lagged cache is computed on the device.  The oracle runs the same sequence function by function
(energy_hydrology.jl:722-906, Runoff.jl:234-283, Simulations.jl:127-135)."""
import os

import numpy as np
import pytest

from helpers import assert_close, cuda_solver, oracle_problem
from test_cuda_host_step import _pin

pytestmark = pytest.mark.gpu
STATE = ("y_theta_l", "y_rho_e_int", "y_theta_i", "y_intf_w", "y_intf_e")


def _problem(ncol, seed):
    import climaland_b200  # noqa: F401
    from climaland_b200 import workloads
    w = workloads.make_workload("energy_hydrology", ncol, N=15, seed=seed, topmodel=True)
    xp = workloads.make_explicit_params(w, seed)
    rng = np.random.default_rng(seed + 100)
    forcing = dict(precip=-rng.uniform(0.0, 4e-7, ncol), f_max=rng.uniform(0.2, 0.6, ncol))
    sat = rng.random(ncol) < 0.35  # a third of the columns with a saturated bottom: runoff terms active
    w["y_theta_l"][sat, :4] = (w["nu"] - w["y_theta_i"])[sat, :4] + 1e-3
    return w, xp, forcing


RUNOFF = dict(f_over=3.28, R_sb=1.484e-4 / 1000, depth=50.0)


def _oracle_step(P, U, p, X, forcing, dt, iters):
    a = P.new_aux()
    P.update_aux(X, U, a)
    dl, di = np.zeros_like(U.theta_l), np.zeros_like(U.theta_l)
    P.phase_change(X, U, a, dl, di)
    R = P.update_runoff(U, forcing["precip"], forcing["f_max"], RUNOFF["f_over"], RUNOFF["R_sb"], RUNOFF["depth"], X=X, a=a)
    for name, v in (("K_lag", a.K), ("kappa_lag", a.kappa), ("theta_l_lag", a.theta_l), ("is_saturated", R.is_saturated),
                    ("R_ss", R.R_ss), ("R_ess", R.R_ess), ("h_grad", R.h_grad)):
        P.set(name, v)
    p.top_bc_w[...] = R.infiltration
    U.theta_l += dt * dl
    U.theta_i += dt * di
    P.implicit_step(U, dt, iters, p=p)
    return float(np.abs(di).max()), float(R.R_ss.max())


def _cuda(w, xp, forcing, out_of_place=False, options=None, **kw):
    from climaland_b200 import workloads
    s = cuda_solver(w, out_of_place=out_of_place, **kw)
    for k, v in (options or {}).items():
        s.set_option(k, v)
    for k, v in xp.items():
        s.set(k, v)
    s.set_explicit_params(**workloads.EXPLICIT_SCALARS)
    s.set("f_max", forcing["f_max"])
    s.set_runoff_params(**RUNOFF)
    return s


def _run(ncol, pinned, steps=1, out_of_place=False, options=None, seed=5):
    dt, iters = 900.0, 3
    w, xp, forcing = _problem(ncol, seed)
    P, U, p = oracle_problem(w, nthreads=os.cpu_count() or 1)
    X = P.explicit_params(**xp)
    s = _cuda(w, xp, forcing, out_of_place, options)
    ins = {k: np.ascontiguousarray(w[k]).copy() for k in STATE}
    ins["precip"] = forcing["precip"].copy()
    pre = "u" if out_of_place else "y"
    outs = {f"{pre}_theta_l": np.zeros((ncol, 15)), f"{pre}_rho_e_int": np.zeros((ncol, 15)), "y_theta_i": np.zeros((ncol, 15)),
            f"{pre}_intf_w": np.zeros(ncol), f"{pre}_intf_e": np.zeros(ncol)}
    if pinned:
        ins = {k: _pin(v) for k, v in ins.items()}
        outs = {k: _pin(v) for k, v in outs.items()}
    moved = []
    for _ in range(steps):
        moved.append(_oracle_step(P, U, p, X, forcing, dt, iters))
        s.soil_step_host(dt, iters, ins, outs)
        for k in STATE:  # the new state is the next step's input, as in a time loop
            ins[k][...] = outs[(pre + k[1:]) if k != "y_theta_i" else k]
    s.close()
    return U, outs, pre, moved


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "pinned"])
@pytest.mark.parametrize("out_of_place", [False, True], ids=["inplace", "outofplace"])
@pytest.mark.parametrize("ncol", [61206, 9001, 700])
def test_soil_step_host_matches_oracle(ncol, out_of_place, pinned):
    U, outs, pre, moved = _run(ncol, pinned, out_of_place=out_of_place)
    assert moved[0][0] > 0.0 and moved[0][1] > 0.0, "the step must freeze or thaw and produce subsurface runoff somewhere"
    # one explicit + implicit step: the per-call bar (1e-12 norm-wise) holds for the whole sequence
    assert_close(outs[f"{pre}_theta_l"], U.theta_l, 1e-12, "theta_l")
    # theta_i = theta_i + dt * (PhaseChange source): where there is (almost) no ice the entry IS dt * source, a difference
    # of operands as large as the field's largest entries (helpers.py: the floor of that kind, 1e-3 of max|b|)
    assert_close(outs["y_theta_i"], U.theta_i, 1e-12, "theta_i", floor_rel=1e-3)
    assert_close(outs[f"{pre}_rho_e_int"], U.rho_e_int, 1e-12, "rho_e_int")
    assert_close(outs[f"{pre}_intf_w"], U.intF_w, 1e-12, "intF_w")
    assert_close(outs[f"{pre}_intf_e"], U.intF_e, 1e-12, "intF_e")


def test_soil_step_host_time_loop():
    """four steps through the chunked route, the outputs fed back as the next inputs"""
    U, outs, pre, _ = _run(20000, True, steps=4)
    assert_close(outs["y_theta_l"], U.theta_l, 1e-11, "theta_l after 4 steps")
    assert_close(outs["y_theta_i"], U.theta_i, 1e-10, "theta_i after 4 steps", floor_rel=1e-3)
    assert_close(outs["y_rho_e_int"], U.rho_e_int, 1e-11, "rho_e_int after 4 steps")


def test_soil_step_host_routes_agree_bitwise():
    res = {}
    for tag, opt, pinned in (("auto", {}, True), ("zerocopy", {"host_route": 3}, True), ("copy_engines", {"host_route": 1}, True),
                             ("plain", {"host_route": 2}, True), ("pageable", {}, False), ("chunks7", {"host_chunks": 7}, True)):
        _, outs, _, _ = _run(20000, pinned, options=opt)
        res[tag] = {k: v.copy() for k, v in outs.items()}
    for tag in ("auto", "zerocopy", "copy_engines", "pageable", "chunks7"):
        for k in res["plain"]:
            assert np.array_equal(res[tag][k], res["plain"][k]), (tag, k)


@pytest.mark.parametrize("form", ["one_launch", "level_fastest", "libm", "percell_kernel", "outofplace"])
def test_resident_soil_step_matches_oracle(form):
    """clb_soil_step: the same whole step on resident mirrors, two steps in a row.  one_launch: column-fastest mirrors,
    FAST arithmetic -- the explicit stage is ONE kernel (cells + per-column sweep through shared memory + in-place
    update); level_fastest: sweep kernel, then cells kernel with the in-place update; libm / percell_kernel: cells
    kernel (source stored), then the sweep, which applies the update."""
    dt, iters, ncol = 900.0, 3, 5003
    out_of_place = form == "outofplace"
    kw = {"level_fastest": dict(layout=2), "libm": dict(math_mode=1)}.get(form, {})
    opts = {"explicit_kernel": 1} if form == "percell_kernel" else None
    w, xp, forcing = _problem(ncol, 9)
    P, U, p = oracle_problem(w, nthreads=os.cpu_count() or 1)
    X = P.explicit_params(**xp)
    s = _cuda(w, xp, forcing, out_of_place, opts, **kw)
    s.set("precip", forcing["precip"])
    for step in range(2):
        a = P.new_aux()
        P.update_aux(X, U, a)  # the cache of the explicit stage at t_n, for the comparison below
        _oracle_step(P, U, p, X, forcing, dt, iters)
        s.soil_step(dt, iters)
        if out_of_place:  # the integrator's U -> u
            for k in ("theta_l", "rho_e_int", "intf_w", "intf_e"):
                s.copy("y_" + k, "u_" + k)
        tol = 1e-12 if step == 0 else 1e-11
        assert_close(s.get("y_theta_l"), U.theta_l, tol, "theta_l")
        assert_close(s.get("y_theta_i"), U.theta_i, tol, "theta_i", floor_rel=1e-3)
        assert_close(s.get("y_rho_e_int"), U.rho_e_int, tol, "rho_e_int")
        assert_close(s.get("y_intf_w"), U.intF_w, tol, "intF_w")
        # what the explicit stage leaves in the cache: update_aux!'s fields and column integrals, the runoff's outputs
        for dev, name in (("p_t", "T"), ("kappa_lag", "kappa"), ("k_lag", "K"), ("total_water", "total_water"),
                          ("total_energy", "total_energy")):
            assert_close(s.get(dev), getattr(a, name), tol, name)
        assert_close(s.get("r_ss"), P.f["R_ss"], tol, "R_ss")
        assert_close(s.get("h_grad"), P.f["h_grad"], tol, "h_grad")
        assert_close(s.get("is_saturated"), P.f["is_saturated"], tol, "is_saturated")
    s.close()


@pytest.mark.parametrize("N", [4, 10, 16, 25, 32])
def test_resident_soil_step_other_level_counts(N):
    """clb_soil_step on column-fastest mirrors at level counts other than the bench's 15: the ONE-launch explicit stage
    (a block is 16 columns x N levels, rounded up to whole warps; the last block of 1 237 columns is ragged) in front of
    whichever stage kernel CLB_VARIANT_AUTO picks for that N"""
    import climaland_b200  # noqa: F401
    from climaland_b200 import workloads
    dt, iters, ncol, seed = 900.0, 3, 1237, 21
    w = workloads.make_workload("energy_hydrology", ncol, N=N, seed=seed, topmodel=True)
    xp = workloads.make_explicit_params(w, seed)
    rng = np.random.default_rng(seed + 100)
    forcing = dict(precip=-rng.uniform(0.0, 4e-7, ncol), f_max=rng.uniform(0.2, 0.6, ncol))
    sat = rng.random(ncol) < 0.35
    w["y_theta_l"][sat, :2] = (w["nu"] - w["y_theta_i"])[sat, :2] + 1e-3
    P, U, p = oracle_problem(w, nthreads=os.cpu_count() or 1)
    X = P.explicit_params(**xp)
    s = _cuda(w, xp, forcing)
    s.set("precip", forcing["precip"])
    a = P.new_aux()
    P.update_aux(X, U, a)
    _oracle_step(P, U, p, X, forcing, dt, iters)
    s.soil_step(dt, iters)
    assert_close(s.get("y_theta_l"), U.theta_l, 1e-12, "theta_l")
    assert_close(s.get("y_theta_i"), U.theta_i, 1e-12, "theta_i", floor_rel=1e-3)
    assert_close(s.get("y_rho_e_int"), U.rho_e_int, 1e-12, "rho_e_int")
    for dev, name in (("p_t", "T"), ("kappa_lag", "kappa"), ("k_lag", "K"), ("total_water", "total_water"),
                      ("total_energy", "total_energy")):
        assert_close(s.get(dev), getattr(a, name), 1e-12, name)
    assert_close(s.get("r_ss"), P.f["R_ss"], 1e-12, "R_ss")
    assert_close(s.get("h_grad"), P.f["h_grad"], 1e-12, "h_grad")
    assert_close(s.get("is_saturated"), P.f["is_saturated"], 1e-12, "is_saturated")
    s.close()


@pytest.mark.parametrize("runoff_model", [0, 1, 2], ids=["norunoff", "surface", "topmodel"])
def test_resident_soil_step_atmos_driven(runoff_model):
    """clb_soil_step with the boundary fluxes of the explicit stage on the device: AtmosDrivenFluxBC top (runoff model +
    assembly from the host's turbulent fluxes / net radiation) and EnergyWaterFreeDrainage bottom; the oracle runs
    update_aux!, the runoff, soil_boundary_fluxes! for both boundaries, PhaseChange, the explicit update, the stage."""
    from test_cuda_explicit_parity import _atmos_forcing
    dt, iters, ncol = 900.0, 3, 3001
    w, xp, forcing = _problem(ncol, 13)
    w["topmodel"] = runoff_model == 2  # the implicit TOPMODELSubsurfaceRunoff source exists with TOPMODELRunoff only
    if runoff_model != 2:
        for k in ("is_saturated", "r_ss", "r_ess", "h_grad"):
            w[k] = np.zeros_like(w[k])
    F = _atmos_forcing(ncol, 3)
    P, U, p = oracle_problem(w, nthreads=os.cpu_count() or 1)
    X = P.explicit_params(**xp)
    s = _cuda(w, xp, forcing, options={"runoff_model": runoff_model, "top_atmos_driven": 1, "bottom_ewfd": 1})
    for k, v in F.items():
        s.set(k, v)
    for step in range(2):
        a = P.new_aux()
        P.update_aux(X, U, a)
        dl, di = np.zeros_like(U.theta_l), np.zeros_like(U.theta_l)
        P.phase_change(X, U, a, dl, di)
        if runoff_model == 2:
            R = P.update_runoff(U, F["precip"], forcing["f_max"], RUNOFF["f_over"], RUNOFF["R_sb"], RUNOFF["depth"], X=X, a=a)
            inf = R.infiltration
            for name, v in (("is_saturated", R.is_saturated), ("R_ss", R.R_ss), ("R_ess", R.R_ess), ("h_grad", R.h_grad)):
                P.set(name, v)
        else:
            _, inf, _ = P.surface_runoff(U, runoff_model, F["precip"], X=X, a=a)
        p.top_bc_w[...], p.top_bc_h[...] = P.atmos_driven_top_fluxes(inf, F["vapor_flux_liq"], F["lhf"], F["shf"], F["r_n"], F["t_air"])
        p.bot_bc_w[...], p.bot_bc_h[...] = P.energy_water_free_drainage(a)
        for name, v in (("K_lag", a.K), ("kappa_lag", a.kappa), ("theta_l_lag", a.theta_l)):
            P.set(name, v)
        U.theta_l += dt * dl
        U.theta_i += dt * di
        P.implicit_step(U, dt, iters, p=p)
        s.soil_step(dt, iters)
        tol = 1e-12 if step == 0 else 1e-11
        assert_close(s.get("top_bc_w"), p.top_bc_w, tol, "top_bc.water")
        assert_close(s.get("top_bc_h"), p.top_bc_h, tol, "top_bc.heat")
        assert_close(s.get("bot_bc_h"), p.bot_bc_h, tol, "bottom_bc.heat")
        assert_close(s.get("y_theta_l"), U.theta_l, tol, "theta_l")
        assert_close(s.get("y_rho_e_int"), U.rho_e_int, tol, "rho_e_int")
        assert_close(s.get("y_intf_e"), U.intF_e, tol, "intF_e")
    s.close()
