#!/usr/bin/env python
"""A whole resident soil step of EnergyHydrology on the ~1 degree column count: update_aux! + PhaseChange,
TOPMODEL runoff, the integrator's explicit update (two axpys, one copy) and the fused implicit stage, all on the
library's mirrors -- no host transfer.  us per step, column-steps/s."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import climaland_b200 as cl  # noqa: F401
from climaland_b200 import workloads
from helpers import cuda_solver

NCOL, N, dt = 61206, 15, 900.0
ss = []
for r in range(4):
    w = workloads.make_workload("energy_hydrology", NCOL, N=N, seed=r, topmodel=True)
    s = cuda_solver(w)
    for k, v in workloads.make_explicit_params(w, r).items():
        s.set(k, v)
    s.set_explicit_params(**workloads.EXPLICIT_SCALARS)
    rng = np.random.default_rng(r)
    s.set("f_max", rng.uniform(0.2, 0.6, NCOL))
    s.set("precip", -rng.uniform(0, 4e-7, NCOL))
    s.set_runoff_params(f_over=3.28, R_sb=1.484e-7, depth=50.0)
    ss.append(s)


def step(s):
    s.set("dye_theta_l", 0.0)
    s.set("dye_theta_i", 0.0)
    s.update_aux_and_phase_change()
    s.update_runoff()
    s.copy("top_bc_w", "infiltration")
    s.axpy("y_theta_l", dt, "dye_theta_l")
    s.axpy("y_theta_i", dt, "dye_theta_i")
    s.implicit_step(dt, 3)


def time_it(fn, label, launches):
    for s in ss:
        fn(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(200):
        fn(ss[k % 4])
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 200
    st = ss[0].implicit_step(dt, 3, want_stats=True)
    print(f"resident EnergyHydrology soil step, {label} ({NCOL} columns x {N} levels, {launches} launches): {us:.1f} us  "
          f"{NCOL / (us * 1e-6):.3e} column-steps/s  SYPD(1 deg) = {dt / (us * 1e-6) / 365:.0f}  nan_count = {st['nan_count']}")


time_it(step, "call by call", 8)
time_it(lambda s: s.soil_step(dt, 3), "clb_soil_step", 2)
for s in ss:
    s.set_option("explicit_kernel", 1)
time_it(lambda s: s.soil_step(dt, 3), "clb_soil_step with the per-cell explicit kernel", 3)
for s in ss:
    s.close()
