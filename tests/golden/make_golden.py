#!/usr/bin/env python
"""Writes tests/golden/stage_vectors.npz: outputs of the CPU oracle (oracle/soil_oracle.c) on small seeded problems,
committed as regression vectors.  The INPUTS are not stored: they are regenerated from the seeds by
climaland_b200.workloads, so a vector pins (workload generator, oracle) together.

These are oracle outputs, not outputs of the reference (Julia, cannot run in the build image): the oracle itself is
pinned on the reference's known-answer tests (reference_kats.json, tests/test_oracle_*.py).  Re-run after an
intentional change of the oracle or of the workload generator:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

NCOL, N = 8, 15
CASES = (("richards", 0, 1800.0, 2), ("energy_hydrology", 0, 900.0, 3), ("energy_hydrology", 1, 900.0, 3))


def problem(model, closure, seed=123):
    import climaland_b200  # noqa: F401
    from climaland_b200 import workloads
    from helpers import oracle_problem
    w = workloads.make_workload(model, NCOL, N=N, seed=seed, topmodel=True)
    if closure == 1:
        from test_cuda_hooks_parity import _to_brooks_corey
        w = _to_brooks_corey(w)
    P, Y, p = oracle_problem(w, closure=closure)
    return w, P, Y, p


def vectors():
    import oracle as orc
    from climaland_b200 import workloads
    out = {}
    for model, closure, dt, iters in CASES:
        tag = f"{model}_cl{closure}"
        w, P, Y, p = problem(model, closure)
        U = Y.copy()
        P.implicit_step(U, dt, iters, p=p)
        out[f"{tag}/stage/theta_l"] = U.theta_l
        out[f"{tag}/stage/intF_w"] = U.intF_w
        if model == "energy_hydrology":
            out[f"{tag}/stage/rho_e_int"] = U.rho_e_int
            out[f"{tag}/stage/intF_e"] = U.intF_e
        P.update_implicit_cache(Y, p)
        dY = P.new_state()
        P.compute_imp_tendency(dY, Y, p)
        W = P.new_jacobian()
        P.compute_jacobian(W, Y, p, dt)
        out[f"{tag}/tendency/theta_l"] = dY.theta_l
        out[f"{tag}/jacobian/w11_di"] = W.w11_di
        if model == "energy_hydrology":
            out[f"{tag}/tendency/rho_e_int"] = dY.rho_e_int
            out[f"{tag}/jacobian/w21_di"] = W.w21_di
            out[f"{tag}/jacobian/w22_up"] = W.w22_up
            if closure == 0:
                xp = workloads.make_explicit_params(w, 123)
                X = P.explicit_params(**xp)
                a = P.new_aux()
                P.update_aux(X, Y, a)
                for k in ("theta_l", "kappa", "T", "K", "psi", "Tf_depressed", "total_water", "total_energy"):
                    out[f"{tag}/aux/{k}"] = getattr(a, k)
                dl, di = np.zeros_like(Y.theta_l), np.zeros_like(Y.theta_l)
                P.phase_change(X, Y, a, dl, di)
                out[f"{tag}/phase_change/dtheta_l"] = dl
                out[f"{tag}/phase_change/dtheta_i"] = di
                rng = np.random.default_rng(5)
                R = P.update_runoff(Y, -rng.uniform(0, 2e-6, NCOL), rng.uniform(0.2, 0.6, NCOL), 3.28, 1.484e-7, 50.0, X=X, a=a)
                for k in ("is_saturated", "h_grad", "infiltration", "R_s", "R_ss", "R_ess"):
                    out[f"{tag}/runoff/{k}"] = getattr(R, k)
    # SoilCO2 stage
    rng = np.random.default_rng(9)
    z_f, z_c = workloads.stretched_grid(N, depth=10.0)
    P = orc.Problem(model=orc.RICHARDS, z_f=z_f, z_c=z_c, ncol=NCOL, nu=0.5, theta_r=0.1, K_sat=1e-6, S_s=1e-3, hcm_a=2.0,
                    hcm_b=2.0, hcm_m=0.5)
    S = P.co2_species(rng.uniform(1e-8, 2e-6, (NCOL, N)), rng.uniform(0.02, 0.45, (NCOL, N)), rng.uniform(1e-4, 4e-4, NCOL))
    C = rng.uniform(5e-5, 2e-3, (NCOL, N))
    top = np.zeros(NCOL)
    P.co2_implicit_step(S, C, top, np.zeros(NCOL), 1800.0, 3)
    out["soilco2/stage/C"] = C
    out["soilco2/stage/top_bc"] = top
    return out


if __name__ == "__main__":
    v = vectors()
    path = os.path.join(HERE, "stage_vectors.npz")
    np.savez_compressed(path, **v)
    print(path, len(v), "vectors", os.path.getsize(path), "bytes")
