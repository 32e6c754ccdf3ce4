#!/usr/bin/env python
"""Times the host-buffer routes of a whole soil step (clb_soil_step_host) on the bench workload with pinned arrays:
CLB_OPT_HOST_ROUTE 0 (the library's choice), 1 (copy engines + staging), 3 (zero-copy kernels) x the chunk count.
python tools/time_host_routes.py -> ms per call (wall clock; the call synchronises)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
import climaland_b200 as cl
from climaland_b200 import workloads
NCOL, dt = 61206, 900.0
w = workloads.make_workload("energy_hydrology", NCOL, N=15, seed=1, topmodel=True)
rng = np.random.default_rng(5)
s = cl.SoilColumnSolver.from_workload(w, out_of_place=True)
for k, v in workloads.make_explicit_params(w, 1).items():
    s.set(k, v)
s.set_explicit_params(**workloads.EXPLICIT_SCALARS)
s.set("f_max", rng.uniform(0.2, 0.6, NCOL))
s.set_runoff_params(f_over=3.28, R_sb=1.484e-7, depth=50.0)


def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    t.numpy()[...] = a
    return t


tin = {k: pinned(w[k]) for k in workloads.STATE}
tin["precip"] = pinned(-rng.uniform(0, 4e-7, NCOL))
tout = {k: pinned(w[k.replace("u_", "y_")]) for k in ("u_theta_l", "u_rho_e_int", "y_theta_i", "u_intf_w", "u_intf_e")}
ins = {a: b.numpy() for a, b in tin.items()}
outs = {a: b.numpy() for a, b in tout.items()}
ref = None
for route in (0, 1, 3):
    for chunks in (4, 8, 12):
        try:
            s.set_option("host_route", route)
        except cl.ClbError:
            continue
        s.set_option("host_chunks", chunks)
        for _ in range(4):
            s.soil_step_host(dt, 3, ins, outs)
        t0 = time.perf_counter()
        n = 40
        for _ in range(n):
            s.soil_step_host(dt, 3, ins, outs)
        ms = 1e3 * (time.perf_counter() - t0) / n
        got = outs["u_theta_l"].copy()
        if ref is None:
            ref = got
        print(f"host_route {route} chunks {chunks:2d}: {ms:.3f} ms per whole soil step  same bits as the first route: {np.array_equal(got, ref)}", flush=True)
s.close()
