"""The committed golden files:
  tests/golden/reference_kats.json  known-answer values the reference's own tests hold (transcribed; file:line inside)
  tests/golden/stage_vectors.npz    outputs of the oracle on small seeded problems (tests/golden/make_golden.py)
CPU: the oracle reproduces both.  GPU (-m gpu): the CUDA library reproduces the vectors through the C ABI."""
import json
import math
import os
import sys

import numpy as np
import pytest

import oracle as orc
from helpers import assert_close, cuda_solver

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
VEC = np.load(os.path.join(HERE, "golden", "stage_vectors.npz"))
L = orc.lib()
RTOL = math.sqrt(np.finfo(np.float64).eps)


def _cmp(got, want, how):
    if how == "equal":
        assert np.all(np.asarray(got, dtype=float) == np.asarray(want, dtype=float)), (got, want)
    else:
        assert np.allclose(got, want, rtol=RTOL, atol=0.0), (got, want)


def test_oracle_reproduces_reference_point_function_kats():
    for k in KATS["point_functions"]:
        how = k["cmp"]
        if k.get("closure") == "vanGenuchten":
            a, n, nu, thr, Ss, Ks = k["alpha"], k["n"], k["nu"], k["theta_r"], k["S_s"], k["K_sat"]
            m = 1 - 1 / n
            _cmp([L.orc_effective_saturation(nu, t, thr) for t in k["theta"]], k["effective_saturation"], how)
            assert L.orc_vg_pressure_head(a, n, m, thr, 0.4, nu, Ss) == k["pressure_head_at_theta_0.4"]
            _cmp(L.orc_vg_pressure_head(a, n, m, thr, 0.5, nu, Ss), k["pressure_head_at_theta_0.5"], how)
            _cmp(L.orc_vg_dpsidtheta(a, n, m, 0.5, nu, thr, Ss), k["dpsidtheta_saturated"], how)
            _cmp(L.orc_vg_hydraulic_conductivity(m, Ks, 1.5), k["hydraulic_conductivity_saturated"], how)
        elif k.get("closure") == "BrooksCorey":
            c, pb, nu, thr, Ss, Ks = k["c"], k["psi_b"], k["nu"], k["theta_r"], k["S_s"], k["K_sat"]
            _cmp(L.orc_bc_matric_potential(c, pb, 1.0), k["matric_potential_at_S_1"], how)
            _cmp(L.orc_bc_pressure_head(c, pb, thr, 0.5, nu, Ss), k["pressure_head_at_theta_0.5"], how)
            _cmp(L.orc_bc_dpsidtheta(c, pb, 0.5, nu, thr, Ss), k["dpsidtheta_saturated"], how)
            _cmp(L.orc_bc_hydraulic_conductivity(c, Ks, 1.5), k["hydraulic_conductivity_saturated"], how)
        elif k["function"] == "impedance_factor":
            _cmp(L.orc_impedance_factor(k["f_i"], k["Omega"]), k["value"], how)
        elif k["function"] == "kappa_sat":
            _cmp(L.orc_kappa_sat(k["theta_l"], k["theta_i"], k["kappa_sat_unfrozen"], k["kappa_sat_frozen"]), k["value"], how)
        elif k["function"] == "relative_saturation":
            _cmp(L.orc_relative_saturation(k["theta_l"], k["theta_i"], k["nu"]), k["value"], how)
        elif k["function"].startswith("EnergyHydrologyParameters"):
            for name in ("alpha", "beta", "gamma", "Omega", "gammaT_ref"):
                assert orc.EXPLICIT_SCALARS[name] == k[name]


def test_oracle_reproduces_reference_jacobian_and_conservation_kats():
    c = KATS["clay"]
    N, dz = c["N"], -c["zmin"] / c["N"]
    m = 1 - 1 / c["vg_n"]
    P = orc.Problem(model=orc.RICHARDS, z_f=np.linspace(c["zmin"], 0.0, N + 1), ncol=1, nu=c["nu"], theta_r=c["theta_r"],
                    K_sat=c["K_sat"], S_s=c["S_s"], hcm_a=c["vg_alpha"], hcm_b=c["vg_n"], hcm_m=m)
    Y, p, W = P.new_state(), P.new_cache(), P.new_jacobian()
    Y.theta_l[...] = c["theta_0"]
    P.update_implicit_cache(Y, p)
    P.compute_jacobian(W, Y, p, c["dtgamma"])
    K = L.orc_vg_hydraulic_conductivity(m, c["K_sat"], L.orc_effective_saturation(c["nu"], c["theta_0"], c["theta_r"]))
    d = L.orc_vg_dpsidtheta(c["vg_alpha"], c["vg_n"], m, c["theta_0"], c["nu"], c["theta_r"], c["S_s"])
    _cmp(W.w11_di[0, 1:-1], -2 * K / dz ** 2 * d - 1, "approx")
    _cmp(W.w11_di[0, [0, -1]], -K / dz ** 2 * d - 1, "approx")
    _cmp(W.w11_up[0, :-1], K / dz ** 2 * d, "approx")
    _cmp(W.w11_lo[0, 1:], K / dz ** 2 * d, "approx")
    k = KATS["conservation"][0]
    p.top_bc_w[...], p.bot_bc_w[...] = k["F_top"], k["F_bot"]
    dY = P.new_state()
    P.compute_imp_tendency(dY, Y, p)
    _cmp(dY.intF_w, k["d_intF"], k["cmp"])


def _oracle_vectors():
    return make_golden.vectors()


def test_oracle_reproduces_the_committed_vectors():
    got = _oracle_vectors()
    assert sorted(got) == sorted(VEC.files)
    for k in VEC.files:
        assert_close(got[k], VEC[k], 1e-13, k)


# --------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("model,closure,dt,iters", make_golden.CASES)
def test_cuda_reproduces_the_committed_vectors(model, closure, dt, iters):
    from climaland_b200 import workloads
    tag = f"{model}_cl{closure}"
    w, P, Y, p = make_golden.problem(model, closure)
    eh = model == "energy_hydrology"
    s = cuda_solver(w, closure=closure)
    s.update_implicit_cache()
    s.compute_imp_tendency()
    s.compute_jacobian(dt)
    assert_close(s.get("dy_theta_l"), VEC[f"{tag}/tendency/theta_l"], 1e-12, "tendency")
    assert_close(s.get("w11_di"), VEC[f"{tag}/jacobian/w11_di"], 1e-12, "w11_di")
    if eh:
        assert_close(s.get("dy_rho_e_int"), VEC[f"{tag}/tendency/rho_e_int"], 1e-12, "tendency rho_e")
        assert_close(s.get("w21_di"), VEC[f"{tag}/jacobian/w21_di"], 1e-12, "w21_di")
        assert_close(s.get("w22_up"), VEC[f"{tag}/jacobian/w22_up"], 1e-12, "w22_up")
        if closure == 0:
            for k, v in workloads.make_explicit_params(w, 123).items():
                s.set(k, v)
            s.set_explicit_params(**workloads.EXPLICIT_SCALARS)
            s.update_aux_and_phase_change()
            for dev, name in (("theta_l_lag", "theta_l"), ("kappa_lag", "kappa"), ("p_t", "T"), ("k_lag", "K"),
                              ("p_psi", "psi"), ("p_tf_depressed", "Tf_depressed"), ("total_water", "total_water"),
                              ("total_energy", "total_energy")):
                assert_close(s.get(dev), VEC[f"{tag}/aux/{name}"], 1e-12, name)
            assert_close(s.get("dye_theta_l"), VEC[f"{tag}/phase_change/dtheta_l"], 1e-12, "phase change")
            rng = np.random.default_rng(5)
            s.set("precip", -rng.uniform(0, 2e-6, make_golden.NCOL))
            s.set("f_max", rng.uniform(0.2, 0.6, make_golden.NCOL))
            s.set_runoff_params(f_over=3.28, R_sb=1.484e-7, depth=50.0)
            s.update_runoff()
            for dev, name in (("h_grad", "h_grad"), ("infiltration", "infiltration"), ("r_ss", "R_ss"), ("r_ess", "R_ess")):
                assert_close(s.get(dev), VEC[f"{tag}/runoff/{name}"], 1e-12, name)
            # update_aux! / update_runoff rewrote the lagged inputs: restore the workload's for the stage below
            for k in ("k_lag", "kappa_lag", "theta_l_lag", "is_saturated", "r_ss", "r_ess", "h_grad"):
                s.set(k, w[k])
    s.implicit_step(dt, iters)
    assert_close(s.get("y_theta_l"), VEC[f"{tag}/stage/theta_l"], 1e-12, "stage theta_l")
    assert_close(s.get("y_intf_w"), VEC[f"{tag}/stage/intF_w"], 1e-12, "stage intF_w")
    if eh:
        assert_close(s.get("y_rho_e_int"), VEC[f"{tag}/stage/rho_e_int"], 1e-12, "stage rho_e_int")
    s.close()
