"""GPU parity of the fine-grained hooks (update_implicit_cache!, compute_imp_tendency!,
compute_jacobian!, ldiv!) against the CPU oracle on identical seeded inputs, through the
C ABI.  Tolerance: 1e-12 norm-wise relative per call (BASELINE.json north_star)."""
import numpy as np
import pytest

from helpers import assert_close, cuda_solver, oracle_problem

pytestmark = pytest.mark.gpu
TOL = 1e-12

import os, sys  # noqa: E401,E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def _workload(model, ncol, N, seed, topmodel=False):
    import climaland_b200  # noqa: F401
    from climaland_b200 import workloads
    return workloads.make_workload(model, ncol, N=N, seed=seed, topmodel=topmodel)


def _to_brooks_corey(w):
    """Brooks-Corey parameters and a state consistent with them: hydrostatic above a water
    table (psi = z_wt - z), S = (psi/psi_b)^(-c), so psi stays within tens of metres.  (Random
    theta with random c gives psi ~ -1e16 m, a problem whose own conditioning is ~1e8.)"""
    rng = np.random.default_rng(7)
    w = dict(w)
    shp = w["nu"].shape
    c = rng.uniform(0.15, 0.6, shp)
    psi_b = -rng.uniform(0.05, 0.5, shp)
    z_wt = rng.uniform(-40.0, -10.0, (shp[0], 1))
    psi = z_wt - w["z_c"][None, :]
    S = np.where(psi < psi_b, (np.minimum(psi, psi_b) / psi_b) ** (-c), 1.0)
    nu_eff = w["nu"] - w.get("y_theta_i", 0.0)
    theta = w["theta_r"] + S * (nu_eff - w["theta_r"])
    theta = theta * (1.0 + rng.uniform(-0.02, 0.02, shp))
    w["y_theta_l"] = np.clip(theta, w["theta_r"] + 1e-3, nu_eff + 5e-3)
    w["hcm_a"], w["hcm_b"] = c, psi_b
    if "theta_l_lag" in w:
        w["theta_l_lag"] = np.minimum(nu_eff, w["y_theta_l"])
    if "is_saturated" in w:
        w["is_saturated"] = (w["y_theta_l"] >= nu_eff).astype(np.float64)
    return w


CASES = [
    # model, closure, top_bc, bottom_bc, topmodel, N, ncol
    ("richards", 0, 0, 0, False, 15, 1000),
    ("richards", 0, 1, 1, False, 15, 333),
    ("richards", 0, 1, 2, True, 20, 130),
    ("richards", 1, 0, 1, False, 15, 257),
    ("richards", 1, 1, 2, True, 7, 65),
    ("energy_hydrology", 0, 0, 0, False, 15, 1000),
    ("energy_hydrology", 0, 0, 0, True, 15, 300),
    ("energy_hydrology", 1, 0, 0, True, 50, 100),
    ("energy_hydrology", 0, 1, 1, False, 15, 64),
]


@pytest.mark.parametrize("layout", [1, 2], ids=["colfast", "levfast"])
@pytest.mark.parametrize("math_mode", [0, 1])
@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}-cl{c[1]}-t{c[2]}b{c[3]}-tm{int(c[4])}-N{c[5]}" for c in CASES])
def test_hooks_match_oracle(case, math_mode, layout):
    model, closure, top_bc, bottom_bc, topmodel, N, ncol = case
    w = _workload(model, ncol, N, seed=11, topmodel=topmodel)
    if closure == 1:
        w = _to_brooks_corey(w)
    rng = np.random.default_rng(3)
    if top_bc == 1:
        w["theta_bc_top"] = w["nu"][:, -1] - rng.uniform(1e-3, 0.1, ncol)
    if bottom_bc == 2:
        w["theta_bc_bot"] = w["nu"][:, 0] - rng.uniform(1e-3, 0.1, ncol)
    P, Y, p = oracle_problem(w, closure, top_bc, bottom_bc)
    s = cuda_solver(w, closure, top_bc, bottom_bc, math_mode=math_mode, layout=layout)
    eh = model == "energy_hydrology"

    # update_implicit_cache!
    P.update_implicit_cache(Y, p)
    s.update_implicit_cache()
    assert_close(s.get("p_psi"), p.psi, TOL, "psi")
    if eh:
        assert_close(s.get("p_t"), p.T, TOL, "T")
    else:
        assert_close(s.get("p_k"), p.K, TOL, "K")
        assert_close(s.get("total_water"), p.total_water, TOL, "total_water")
        if top_bc == 1:
            assert_close(s.get("top_bc_w"), p.top_bc_w, TOL, "top_bc")
            assert_close(s.get("bot_bc_w"), p.bot_bc_w, TOL, "bot_bc")
            assert_close(s.get("dfluxbcdy"), p.dfluxBCdY, TOL, "dfluxBCdY")

    # compute_imp_tendency!
    dY = P.new_state()
    P.compute_imp_tendency(dY, Y, p)
    s.compute_imp_tendency()
    assert_close(s.get("dy_theta_l"), dY.theta_l, TOL, "dY.theta_l")
    assert_close(s.get("dy_intf_w"), dY.intF_w, TOL, "dY.intF_w")
    if eh:
        assert_close(s.get("dy_rho_e_int"), dY.rho_e_int, TOL, "dY.rho_e_int")
        assert_close(s.get("dy_intf_e"), dY.intF_e, TOL, "dY.intF_e")
        assert np.all(s.get("dy_theta_i") == 0.0)

    # compute_jacobian!
    dtg = 900.0
    W = P.new_jacobian()
    P.compute_jacobian(W, Y, p, dtg)
    s.compute_jacobian(dtg)
    blocks = ["w11"] + (["w21", "w22"] if eh else [])
    for b in blocks:
        for d in ("lo", "di", "up"):
            assert_close(s.get(f"{b}_{d}"), getattr(W, f"{b}_{d}"), TOL, f"{b}_{d}")

    # ldiv!
    b = P.new_state()
    b.theta_l[...] = rng.normal(0, 1e-3, b.theta_l.shape)
    b.intF_w[...] = rng.normal(0, 1.0, ncol)
    s.set("b_theta_l", b.theta_l)
    s.set("b_intf_w", b.intF_w)
    if eh:
        b.rho_e_int[...] = rng.normal(0, 1e4, b.theta_l.shape)
        b.theta_i[...] = rng.normal(0, 1e-3, b.theta_l.shape)
        b.intF_e[...] = rng.normal(0, 1.0, ncol)
        s.set("b_rho_e_int", b.rho_e_int)
        s.set("b_theta_i", b.theta_i)
        s.set("b_intf_e", b.intF_e)
    x = P.new_state()
    P.ldiv(x, W, b)
    s.ldiv()
    assert_close(s.get("x_theta_l"), x.theta_l, TOL, "x.theta_l")
    assert_close(s.get("x_intf_w"), x.intF_w, TOL, "x.intF_w")
    if eh:
        assert_close(s.get("x_rho_e_int"), x.rho_e_int, TOL, "x.rho_e_int")
        assert_close(s.get("x_theta_i"), x.theta_i, TOL, "x.theta_i")
        assert_close(s.get("x_intf_e"), x.intF_e, TOL, "x.intF_e")
    s.close()


def test_diagonal_block_solve_and_resident_vector_ops():
    """clb_ldiv_diagonal (DiagonalMatrixRow blocks of surface variables, implicit_timestepping.jl:117-121; the canopy
    temperature block of canopy_energy.jl:222-250), clb_field_axpy and clb_field_copy against numpy"""
    import climaland_b200 as cl
    w = _workload("richards", 777, 15, seed=9)
    s = cuda_solver(w)
    rng = np.random.default_rng(1)
    dtg = 450.0
    # the canopy block: dtgamma (dLW_n/dT - dshf/dT - dlhf/dT) / (ac_canopy max(LAI, eps)) - 1
    wd = dtg * (-rng.uniform(1.0, 8.0, 777) - rng.uniform(5, 30, 777) - rng.uniform(0, 40, 777)) / (2e3 * np.maximum(rng.uniform(0, 6, 777), 1e-16)) - 1.0
    b = rng.normal(0, 1.0, 777)
    s.set("sfc_w_di", wd)
    s.set("sfc_b", b)
    s.ldiv_diagonal("sfc_w_di", "sfc_b", "sfc_x")
    assert np.array_equal(s.get("sfc_x"), b / wd)
    y0, x0 = s.get("y_theta_l"), rng.normal(0, 1e-7, (777, 15))
    s.set("dy_theta_l", x0)
    s.axpy("y_theta_l", 1800.0, "dy_theta_l")
    assert np.array_equal(s.get("y_theta_l"), y0 + 1800.0 * x0)
    s.copy("x_theta_l", "y_theta_l")
    assert np.array_equal(s.get("x_theta_l"), s.get("y_theta_l"))
    with pytest.raises(cl.ClbError):
        s.ldiv_diagonal("sfc_w_di", "y_theta_l", "sfc_x")   # mixed kinds
    s.close()


@pytest.mark.parametrize("layout", [1, 2], ids=["colfast", "levfast"])
def test_ldiv_all_blocks_of_an_integrated_model(layout):
    """clb_ldiv_all: the ldiv! of an integrated model's FieldMatrixWithSolver (implicit_timestepping.jl:63-172) in one
    launch -- EnergyHydrology's blocks (BlockLowerTriangularSolve(theta_l)) against the oracle's ldiv, the SoilCO2
    tridiagonals against scipy's banded solver on the rows clb_soilco2_compute_jacobian left in the mirrors, the canopy
    temperature's DiagonalMatrixRow block against b / w; and block by block equal to the separate entry points."""
    from scipy.linalg import solve_banded
    ncol, N, dtg = 513, 15, 900.0
    w = _workload("energy_hydrology", ncol, N, seed=19, topmodel=True)
    rng = np.random.default_rng(3)
    P, Y, p = oracle_problem(w)
    s = cuda_solver(w, layout=layout)
    P.update_implicit_cache(Y, p)
    W = P.new_jacobian()
    P.compute_jacobian(W, Y, p, dtg)
    s.update_implicit_cache()
    s.compute_jacobian(dtg)
    b = P.new_state()
    b.theta_l[...] = rng.normal(0, 1e-3, b.theta_l.shape)
    b.rho_e_int[...] = rng.normal(0, 1e4, b.theta_l.shape)
    b.theta_i[...] = rng.normal(0, 1e-3, b.theta_l.shape)
    b.intF_w[...], b.intF_e[...] = rng.normal(0, 1.0, ncol), rng.normal(0, 1.0, ncol)
    for k, v in (("b_theta_l", b.theta_l), ("b_rho_e_int", b.rho_e_int), ("b_theta_i", b.theta_i), ("b_intf_w", b.intF_w),
                 ("b_intf_e", b.intF_e)):
        s.set(k, v)
    # SoilCO2 blocks
    rhs = {}
    for name in ("co2", "o2"):
        s.set(f"{name}_y", rng.uniform(5e-5, 2e-3, (ncol, N)))
        s.set(f"{name}_d", rng.uniform(1e-8, 2e-6, (ncol, N)))
        s.set(f"{name}_theta_eff", rng.uniform(0.02, 0.45, (ncol, N)))
        s.set(f"{name}_bot_bc", np.zeros(ncol))
        s.set(f"{name}_top_bc", np.zeros(ncol))
        rhs[name] = rng.normal(0, 1e-4, (ncol, N))
        s.set(f"{name}_b", rhs[name])
    s.set_co2_top_state(co2=False, o2=False)
    s.soilco2_compute_jacobian(dtg)
    # canopy temperature block
    wd = dtg * (-rng.uniform(6.0, 80.0, ncol)) / (2e3 * np.maximum(rng.uniform(0, 6, ncol), 1e-16)) - 1.0
    bs = rng.normal(0, 1.0, ncol)
    s.set("sfc_w_di", wd)
    s.set("sfc_b", bs)
    s.ldiv_all(soil=True, soilco2=True, surface=True)
    x = P.new_state()
    P.ldiv(x, W, b)
    assert_close(s.get("x_theta_l"), x.theta_l, TOL, "x.theta_l")
    assert_close(s.get("x_rho_e_int"), x.rho_e_int, TOL, "x.rho_e_int")
    assert_close(s.get("x_theta_i"), x.theta_i, TOL, "x.theta_i")
    assert_close(s.get("x_intf_w"), x.intF_w, TOL, "x.intF_w")
    assert_close(s.get("x_intf_e"), x.intF_e, TOL, "x.intF_e")
    for name in ("co2", "o2"):
        lo, di, up = s.get(f"{name}_w_lo"), s.get(f"{name}_w_di"), s.get(f"{name}_w_up")
        got = s.get(f"{name}_x")
        for c in range(0, ncol, 37):
            ab = np.zeros((3, N))
            ab[0, 1:], ab[1], ab[2, :-1] = up[c, :-1], di[c], lo[c, 1:]
            assert_close(got[c], solve_banded((1, 1), ab, rhs[name][c]), 1e-11, f"x.{name}")
    assert np.array_equal(s.get("sfc_x"), bs / wd)
    # the separate entry points give the same bits
    xs = {k: s.get(k) for k in ("x_theta_l", "x_rho_e_int", "sfc_x")}
    s.ldiv()
    s.ldiv_diagonal("sfc_w_di", "sfc_b", "sfc_x")
    for k, v in xs.items():
        assert np.array_equal(s.get(k), v), k
    import climaland_b200 as cl
    with pytest.raises(cl.ClbError):
        s.L.clb_ldiv_all  # noqa: B018 (exists)
        s.ldiv_all(soil=False, soilco2=False, surface=False)
    s.close()
