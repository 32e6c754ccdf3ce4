"""Test plumbing: build the CPU oracle problem and the CUDA solver from the same
workload dict, and compare fields.  The oracle is the checker only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle as orc  # noqa: E402

ORC_FIELD = dict(nu="nu", theta_r="theta_r", K_sat="K_sat", S_s="S_s", hcm_a="hcm_a", hcm_b="hcm_b",
                 hcm_m="hcm_m", rho_c_ds="rho_c_ds", k_lag="K_lag", kappa_lag="kappa_lag",
                 theta_l_lag="theta_l_lag", is_saturated="is_saturated", r_ss="R_ss", r_ess="R_ess",
                 h_grad="h_grad", theta_bc_top="theta_bc_top", theta_bc_bot="theta_bc_bot")
STATE = dict(y_theta_l="theta_l", y_rho_e_int="rho_e_int", y_theta_i="theta_i", y_intf_w="intF_w",
             y_intf_e="intF_e")
CACHE_BC = dict(top_bc_w="top_bc_w", bot_bc_w="bot_bc_w", top_bc_h="top_bc_h", bot_bc_h="bot_bc_h")


def oracle_problem(w, closure=0, top_bc=0, bottom_bc=0, nthreads=1):
    model = orc.RICHARDS if w["model"] == "richards" else orc.ENERGY_HYDROLOGY
    fields = {ORC_FIELD[k]: v for k, v in w.items() if k in ORC_FIELD}
    P = orc.Problem(model=model, closure=closure, top_bc=top_bc, bottom_bc=bottom_bc,
                    has_topmodel_source=w.get("topmodel", False), z_f=w["z_f"], z_c=w["z_c"], ncol=w["ncol"],
                    nthreads=nthreads, **fields)
    Y, p = P.new_state(), P.new_cache()
    for k, a in STATE.items():
        if k in w:
            getattr(Y, a)[...] = w[k]
    for k, a in CACHE_BC.items():
        if k in w:
            getattr(p, a)[...] = w[k]
    return P, Y, p


def cuda_solver(w, closure=0, top_bc=0, bottom_bc=0, **kw):
    import climaland_b200 as cl
    return cl.SoilColumnSolver.from_workload(w, closure=closure, top_bc=top_bc, bottom_bc=bottom_bc, **kw)


def rel_err(a, b):
    """max|a-b| / max|b| (norm-wise relative error; 0 if both are all-zero)"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    num = np.max(np.abs(a - b)) if a.size else 0.0
    if den == 0.0:
        return num
    return num / den


def assert_close(a, b, tol, what=""):
    assert np.all(np.isfinite(a)) == np.all(np.isfinite(b)), f"{what}: finiteness differs"
    e = rel_err(a, b)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"
