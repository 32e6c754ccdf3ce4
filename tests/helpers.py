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


def elem_rel_err(a, b, floor_rel=1e-6, floor_abs=0.0):
    """max_i |a_i - b_i| / max(|b_i|, floor): ELEMENT-wise relative error.  The floor keeps entries that are zero or
    cancel to (almost) zero from dividing by nothing: floor = max(floor_abs, floor_rel * max|b|), i.e. by default every
    entry down to six decades below the field's largest magnitude is held to the full relative tolerance."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    floor = max(floor_abs, floor_rel * float(np.max(np.abs(b))))
    if floor == 0.0:
        return float(np.max(np.abs(a - b)))
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


# Floors of the element-wise check, by the KIND of quantity (matched on the label the test passes):
#   state-like (theta, K, T, kappa, Jacobian entries, totals, parameters): 1e-6 of the field's largest magnitude -- every
#       entry within six decades of the maximum is held to the full relative tolerance;
#   differences of fluxes (tendencies dY.*, sources, phase change, state-type boundary fluxes K (psi_bc - psi)/dz, Newton
#       increments x = W^-1 f, the flux integrals intF = temp + dtgamma (F_bot - F_top)): an entry is a difference of
#       operands as large as the field's largest entries and carries THEIR rounding, so it is held relative to
#       max(|b_i|, 1e-3 max|b|);
#   rho_e_int: energy relative to T_ref, crossing zero; the reference itself resolves it only to rho_c ulp(T_ref) =
#       2.5e6 x 5.7e-14 = 1.4e-7 J/m3 because T = T_ref + rho_e/rho_c is formed as an absolute temperature
#       (energy_hydrology.jl:427-445), i.e. 1e-12 relative only above ~1.4e5 J/m3 = 2e-3 of max|rho_e|: floor 1e-2;
#   Jacobian entries w11 / w21 / w22: products of TWO closure outputs (a face mean of K and dpsi/dtheta), each held to
#       the tolerance on its own: element-wise tolerance 2x the norm-wise one;
#   psi: the van Genuchten matric potential goes to 0 as S -> 1 and loses digits there in ANY double evaluation of the
#       reference's formula (S^(-1/m) - 1 cancels: relative error eps / (1 - S)); it only ever enters as psi + z with
#       |z| >= 0.025 m, so it is held relative to max(|psi_i|, 0.01 m), element-wise tolerance 10x the norm-wise one.
_DIFFERENCE = ("dy", "tendency", "source", "phase change", "intf", "top_bc", "bot_bc", "x.", "x_", "dflux")
_ENERGY = ("rho_e",)


def kind_of(what):
    w = what.lower()
    if any(k in w for k in _DIFFERENCE):
        return "difference", 1e-3, 1.0, 0.0
    if any(k in w for k in _ENERGY):
        return "energy", 1e-2, 1.0, 0.0
    if w.startswith("psi") or w == "p_psi":
        return "psi", 1e-6, 10.0, 1e-2
    if w[:3] in ("w11", "w21", "w22"):
        return "jacobian", 1e-6, 2.0, 0.0
    return "state", 1e-6, 1.0, 0.0


def assert_close(a, b, tol, what="", elem_tol=None, floor_rel=None, floor_abs=None):
    """Two metrics, both must hold: the norm-wise relative error max|a-b| / max|b| <= tol and the element-wise
    relative error (elem_rel_err) <= elem_tol, with the floor of the quantity's kind (table above) unless the caller
    states another."""
    assert np.all(np.isfinite(a)) == np.all(np.isfinite(b)), f"{what}: finiteness differs"
    e = rel_err(a, b)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"
    kind, fl, mult, fa = kind_of(what)
    fl = fl if floor_rel is None else floor_rel
    floor_abs = fa if floor_abs is None else floor_abs
    et = tol * mult if elem_tol is None else elem_tol
    ee = elem_rel_err(a, b, fl, floor_abs)
    assert ee <= et, f"{what} [{kind}]: element-wise relative error {ee:.3e} > {et:.1e} (floor {fl:.0e} of max|b|)"
