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
]


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


@pytest.mark.parametrize("variant", [0, 2], ids=["auto", "generic"])
@pytest.mark.parametrize("math_mode", [0, 1], ids=["fast", "libm"])
@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}-cl{c[1]}-t{c[2]}b{c[3]}-tm{int(c[4])}-N{c[5]}-it{c[8]}" for c in CASES])
def test_fused_step_matches_oracle(case, math_mode, variant):
    model, closure, top_bc, bottom_bc, topmodel, N, ncol, dt, iters = case
    w = _setup(case)
    P, U, p = oracle_problem(w, closure, top_bc, bottom_bc)
    s = cuda_solver(w, closure, top_bc, bottom_bc, math_mode=math_mode, kernel_variant=variant)
    it, nrm = P.implicit_step(U, dt, iters, tol=-1.0, p=p)
    st = s.implicit_step(dt, iters, want_stats=True)
    assert st["iterations"] == iters and st["nan_count"] == 0
    _compare_state(s, U, model == "energy_hydrology")
    assert abs(st["dx_norm"] - nrm) <= 1e-9 * max(nrm, 1e-300)
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


def test_repeated_steps_are_deterministic():
    w = _setup(("energy_hydrology", 0, 0, 0, True, 15, 2000, 900.0, 3))
    out = []
    for _ in range(2):
        s = cuda_solver(w)
        s.implicit_step(900.0, 3)
        out.append((s.get("y_theta_l"), s.get("y_rho_e_int")))
        s.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
