"""GPU parity of the fused implicit ARS111 stage (clb_implicit_step) against the CPU
oracle's Newton loop, through the C ABI: register-column and generic variants, fixed
iterations and the tolerance path."""
import numpy as np
import pytest

from helpers import assert_close, cuda_solver, oracle_problem
from test_cuda_hooks_parity import _to_brooks_corey, _workload

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = [
    # model, closure, top_bc, bottom_bc, topmodel, N, ncol, dt, iters
    ("richards", 0, 0, 0, False, 15, 1000, 1800.0, 2),
    ("richards", 0, 1, 1, False, 15, 333, 1800.0, 3),
    ("richards", 0, 1, 2, True, 15, 130, 450.0, 2),
    ("richards", 1, 0, 1, True, 15, 257, 1800.0, 2),
    ("richards", 0, 0, 0, True, 50, 100, 1800.0, 2),
    ("energy_hydrology", 0, 0, 0, True, 15, 1000, 900.0, 3),
    ("energy_hydrology", 0, 0, 0, False, 15, 300, 900.0, 1),
    ("energy_hydrology", 1, 0, 0, True, 15, 100, 900.0, 3),
    ("energy_hydrology", 0, 0, 0, True, 50, 100, 900.0, 3),
    ("richards", 0, 1, 1, True, 24, 77, 1800.0, 2),
    ("energy_hydrology", 0, 0, 0, True, 30, 90, 900.0, 3),
    ("energy_hydrology", 1, 0, 0, False, 9, 45, 900.0, 2),
    ("energy_hydrology", 0, 0, 0, True, 16, 211, 900.0, 3),
    ("richards", 0, 0, 1, True, 16, 97, 1800.0, 2),
    ("energy_hydrology", 0, 0, 0, True, 15, 16, 900.0, 3),
    ("energy_hydrology", 0, 0, 0, True, 15, 1, 900.0, 3),
    ("richards", 1, 0, 0, True, 50, 203, 1800.0, 2),
    ("energy_hydrology", 1, 0, 0, False, 50, 61, 900.0, 3),
    ("richards", 0, 0, 0, False, 50, 3, 1800.0, 2),
    # level counts the octet takes at run time (8 lanes x Q cells, NR = 8 Q rows); 33 / 17 / 25: the pads at the top
    # span whole lanes, so the top boundary enters in an inner part
    ("richards", 0, 0, 0, True, 20, 130, 1800.0, 2),
    ("energy_hydrology", 0, 0, 0, True, 40, 70, 900.0, 3),
    ("richards", 1, 0, 1, True, 33, 61, 1800.0, 2),
    ("energy_hydrology", 0, 0, 0, False, 17, 45, 900.0, 3),
    ("energy_hydrology", 1, 0, 0, True, 25, 38, 900.0, 2),
    ("richards", 0, 0, 0, False, 48, 9, 1800.0, 2),
    ("energy_hydrology", 0, 0, 0, True, 41, 21, 900.0, 3),
    # 49 .. 64 levels (other than the compiled N = 50): Q = 7 / 8 cells per lane, single-buffered tiles
    ("richards", 0, 0, 0, True, 56, 37, 1800.0, 2),
    ("energy_hydrology", 0, 0, 0, True, 49, 29, 900.0, 3),
    ("richards", 1, 0, 1, False, 57, 21, 1800.0, 2),
    ("energy_hydrology", 1, 0, 0, True, 64, 13, 900.0, 3),
    ("energy_hydrology", 0, 0, 0, False, 60, 50, 900.0, 2),
    # MoistureStateBC top (boundary fluxes and dfluxBCdY at the iterate): Brooks-Corey, N = 16 (no pad row: the top cell
    # is slot 0 of its lane), a lagged bottom flux value, one column more than a tile
    ("richards", 1, 1, 1, False, 15, 90, 1800.0, 3),
    ("richards", 0, 1, 0, True, 16, 41, 900.0, 2),
    ("richards", 1, 1, 2, True, 16, 9, 1800.0, 3),
]

# (kernel_variant, layout): lane-per-cell on level-fastest mirrors, register-column and generic on
# column-fastest mirrors, generic on level-fastest mirrors
VARIANTS = {"lane_per_cell": (3, 2), "register_column": (1, 1), "generic_cf": (2, 1), "generic_lf": (2, 2),
            "lane_quad": (4, 1), "lane_quad_pipelined": (5, 1), "lane_octet": (6, 1), "lane_octet_lf": (6, 2), "auto": (0, 0)}


def _setup(case):
    model, closure, top_bc, bottom_bc, topmodel, N, ncol, dt, iters = case
    w = _workload(model, ncol, N, seed=5, topmodel=topmodel)
    if closure == 1:
        w = _to_brooks_corey(w)
    rng = np.random.default_rng(3)
    if top_bc == 1:
        w["theta_bc_top"] = w["nu"][:, -1] - rng.uniform(1e-3, 0.1, ncol)
    if bottom_bc == 2:
        w["theta_bc_bot"] = w["nu"][:, 0] - rng.uniform(1e-3, 0.1, ncol)
    w["y_intf_w"] = rng.normal(0, 1e-3, ncol)
    if model == "energy_hydrology":
        w["y_intf_e"] = rng.normal(0, 1e3, ncol)
    return w


def _compare_state(s, U, eh, tol=TOL):
    assert_close(s.get("y_theta_l"), U.theta_l, tol, "theta_l")
    assert_close(s.get("y_intf_w"), U.intF_w, tol, "intF_w")
    if eh:
        assert_close(s.get("y_rho_e_int"), U.rho_e_int, tol, "rho_e_int")
        assert_close(s.get("y_intf_e"), U.intF_e, tol, "intF_e")
        assert_close(s.get("y_theta_i"), U.theta_i, 0.0, "theta_i")


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("math_mode", [0, 1], ids=["fast", "libm"])
@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}-cl{c[1]}-t{c[2]}b{c[3]}-tm{int(c[4])}-N{c[5]}-it{c[8]}" for c in CASES])
def test_fused_step_matches_oracle(case, math_mode, variant):
    model, closure, top_bc, bottom_bc, topmodel, N, ncol, dt, iters = case
    kv, layout = VARIANTS[variant]
    if variant == "lane_per_cell" and N > 31:
        pytest.skip("lane-per-cell needs N <= 31")
    if variant == "register_column" and N != 15:
        pytest.skip("register-column is built for N = 15")
    if variant.startswith("lane_quad") and (N not in (15, 16) or math_mode != 0 or
                                            (model == "richards" and top_bc == 1 and variant != "lane_quad_pipelined")):
        pytest.skip("lane-quad is built for N = 15 / 16, fast math, column-fastest mirrors; a MoistureStateBC top "
                    "(boundary fluxes at the iterate): the pipelined quad only")
    if variant.startswith("lane_octet") and (not (15 <= N <= 64) or math_mode != 0
                                             or (model == "richards" and top_bc == 1)):
        pytest.skip("lane-octet is built for N = 15 .. 64, fast math, flux boundary conditions")
    if variant == "lane_octet_lf" and N != 50:
        pytest.skip("only the N = 50 octet reads level-fastest mirrors")
    w = _setup(case)
    P, U, p = oracle_problem(w, closure, top_bc, bottom_bc)
    s = cuda_solver(w, closure, top_bc, bottom_bc, math_mode=math_mode, kernel_variant=kv, layout=layout)
    it, nrm = P.implicit_step(U, dt, iters, tol=-1.0, p=p)
    st = s.implicit_step(dt, iters, want_stats=True)
    assert st["iterations"] == iters and st["nan_count"] == 0
    _compare_state(s, U, model == "energy_hydrology")
    assert abs(st["dx_norm"] - nrm) <= 1e-9 * max(nrm, 1e-300)
    if model == "richards" and top_bc == 1:
        # update_implicit_boundary_fluxes (rre.jl:460-468) leaves its last evaluation in p.soil.top_bc / bottom_bc
        assert_close(s.get("top_bc_w"), p.top_bc_w, 1e-11, "top_bc")
        assert_close(s.get("bot_bc_w"), p.bot_bc_w, 1e-11, "bot_bc")
    # the step really moved the state
    assert np.max(np.abs(s.get("y_theta_l") - w["y_theta_l"])) > 0
    s.close()


@pytest.mark.parametrize("model", ["richards", "energy_hydrology"])
def test_tolerance_path_matches_oracle(model):
    """ConvergenceChecker-style stopping (experiments/standalone/Soil/richards_comparison.jl:77-86):
    the norm is taken over all columns; iterations stop without a host round trip."""
    case = (model, 0, 0, 0, False, 15, 500, 900.0, 20)
    w = _setup(case)
    # pick a tolerance well inside a gap of the oracle's own norm sequence
    norms = []
    for k in range(1, 9):
        P, U, p = oracle_problem(w)
        norms.append(P.implicit_step(U, 900.0, k, p=p)[1])
    k = next(k for k in range(2, 8) if norms[k] < 0.7 * min(norms[:k]))
    tol = float(np.sqrt(norms[k] * min(norms[:k])))
    P, U, p = oracle_problem(w)
    s = cuda_solver(w)
    it, nrm = P.implicit_step(U, 900.0, 20, tol=tol, p=p)
    st = s.implicit_step(900.0, 20, tol=tol, want_stats=True)
    assert it == k + 1
    assert 1 < it < 20, "pick a tolerance that stops the oracle early"
    assert st["iterations"] == it and st["converged"]
    assert abs(st["dx_norm"] - nrm) <= 1e-6 * nrm
    _compare_state(s, U, model == "energy_hydrology", tol=1e-11)
    s.close()


@pytest.mark.parametrize("model", ["richards", "energy_hydrology"])
def test_out_of_place_leaves_temp_untouched(model):
    """CLB_OPT_OUT_OF_PLACE: Y keeps ClimaTimeSteppers' `temp`, the new stage value goes to U."""
    case = (model, 0, 0, 0, True, 15, 300, 900.0, 3)
    w = _setup(case)
    P, U, p = oracle_problem(w)
    P.implicit_step(U, 900.0, 3, p=p)
    s = cuda_solver(w, out_of_place=True)
    for _ in range(2):  # idempotent: the second call repeats the first
        s.implicit_step(900.0, 3)
        assert np.array_equal(s.get("y_theta_l"), w["y_theta_l"])
        assert_close(s.get("u_theta_l"), U.theta_l, TOL, "u_theta_l")
        assert_close(s.get("u_intf_w"), U.intF_w, TOL, "u_intf_w")
        if model == "energy_hydrology":
            assert np.array_equal(s.get("y_rho_e_int"), w["y_rho_e_int"])
            assert_close(s.get("u_rho_e_int"), U.rho_e_int, TOL, "u_rho_e_int")
            assert_close(s.get("u_intf_e"), U.intF_e, TOL, "u_intf_e")
    s.close()


def test_repeated_steps_are_deterministic():
    w = _setup(("energy_hydrology", 0, 0, 0, True, 15, 2000, 900.0, 3))
    out = []
    for _ in range(2):
        s = cuda_solver(w)
        s.implicit_step(900.0, 3)
        out.append((s.get("y_theta_l"), s.get("y_rho_e_int")))
        s.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
