#!/usr/bin/env python
"""The fine-grained entry points (the per-call drop-in level of INTEGRATION.md: cache_imp!, Wfact, T_imp!, ldiv!) timed
one by one on the bench workload, and the Newton stage they add up to when ClimaTimeSteppers drives them call by call
(max_iters x (cache_imp! + Wfact + T_imp! + ldiv! + two axpys)), beside the fused stage.  us per call, GB/s of each
call's own minimum traffic (fields read once + fields written once)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import climaland_b200 as cl
from climaland_b200 import workloads

NCOL, N = 61206, 15
args = [a for a in sys.argv[1:]]
models = args or ["energy_hydrology", "richards"]
for model in models:
    iters, dt = (3, 900.0) if model == "energy_hydrology" else (2, 1800.0)
    eh = model == "energy_hydrology"
    w = workloads.make_workload(model, NCOL, N=N, seed=1, topmodel=True)
    ss = [cl.SoilColumnSolver.from_workload(w) for _ in range(4)]  # 4 field sets: more than the L2 holds
    for s in ss:
        s.update_implicit_cache(); s.compute_jacobian(dt); s.compute_imp_tendency()
        for k in (("theta_l", "rho_e_int") if eh else ("theta_l",)):
            s.copy("b_" + k, "dy_" + k)
        s.ldiv()
    torch.cuda.synchronize()

    def time_it(fn, reps=200):
        for s in ss: fn(s)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(reps): fn(ss[k % 4])
            e1.record(); torch.cuda.synchronize()
            best = min(best, 1e3 * e0.elapsed_time(e1) / reps)
        return best

    cells = NCOL * N * 8
    # minimum traffic per call in cell fields (per-column fields left out): read + written
    nparam = 7  # nu, theta_r, K_sat, S_s, alpha, n, m
    traffic = {
        "update_implicit_cache": ((nparam + 1 + (3 if eh else 0)) + (2 if eh else 2)),          # state + params -> psi, T (K, psi)
        "compute_jacobian": ((nparam + 1 + (5 if eh else 1)) + (9 if eh else 3)),              # + lagged K, kappa, theta_l ... -> W rows
        "compute_imp_tendency": ((4 if eh else 2) + 1 + (2 if eh else 1)),                     # K, psi (kappa, T), is_saturated -> dY
        "ldiv": ((9 if eh else 3) + (2 if eh else 1) + (2 if eh else 1)),                      # W rows, b -> x
    }
    total = 0.0
    print(f"{model}: {NCOL} columns x {N} levels")
    for name, fn in (("update_implicit_cache", lambda s: s.update_implicit_cache()),
                     ("compute_jacobian", lambda s: s.compute_jacobian(dt)),
                     ("compute_imp_tendency", lambda s: s.compute_imp_tendency()),
                     ("ldiv", lambda s: s.ldiv())):
        us = time_it(fn)
        total += us
        gb = traffic[name] * cells / (us * 1e-6) / 1e9
        print(f"  {name:24s} {us:8.1f} us   {traffic[name]:2d} cell fields  {gb:7.0f} GB/s", flush=True)
    y = "y_theta_l"
    us_axpy = time_it(lambda s: s.axpy(y, -1.0, "x_theta_l"))
    nax = 2 if eh else 1
    stage = iters * (total + nax * us_axpy)
    us_fused = time_it(lambda s: s.implicit_step(dt, iters))
    print(f"  axpy (one field)         {us_axpy:8.1f} us")
    print(f"  call-by-call Newton stage = {iters} x (the four calls + {nax} axpy): {stage:8.1f} us;  fused stage (one launch): {us_fused:6.1f} us"
          f"  ({stage / us_fused:.1f}x)", flush=True)
    for s in ss: s.close()
