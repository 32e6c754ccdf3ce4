#!/usr/bin/env python
"""Times the fused SoilCO2 implicit stage (clb_soilco2_implicit_step: CO2 and O2 tridiagonals, 3 Newton
iterations) on the ~1 degree column count: us per stage, column-steps/s, algorithmic GB/s and HBM fraction."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import climaland_b200 as cl
from climaland_b200 import workloads

peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
for ncol, N in ((61206, 15), (1_000_000, 15), (100_000, 50)):
    rng = np.random.default_rng(0)
    z_f, z_c = workloads.stretched_grid(N)
    ss = []
    for r in range(4 if ncol < 500_000 else 1):
        s = cl.SoilColumnSolver(model=cl.RICHARDS, n_columns=ncol, z_f=z_f, z_c=z_c)
        for name in ("co2", "o2"):
            s.set(f"{name}_y", rng.uniform(5e-5, 2e-3, (ncol, N)))
            s.set(f"{name}_d", rng.uniform(1e-8, 2e-6, (ncol, N)))
            s.set(f"{name}_theta_eff", rng.uniform(0.02, 0.45, (ncol, N)))
            s.set(f"{name}_c_atm", rng.uniform(1e-4, 4e-4, ncol))
        s.set_co2_top_state(co2=True, o2=True)
        ss.append(s)
    for s in ss:
        s.soilco2_implicit_step(1800.0, 3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 200 if ncol < 500_000 else 30
    e0.record()
    for k in range(reps):
        ss[k % len(ss)].soilco2_implicit_step(1800.0, 3)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    bytes_cs = 8 * (2 * 4 * N + 6)   # per species: C, D, theta_eff in, C out; top_bc, c_atm, bottom_bc per column
    gbs = ncol * bytes_cs / (us * 1e-6) / 1e9
    print(f"soilco2 stage  {ncol:8d} columns x {N:2d} levels: {us:8.1f} us  {ncol / (us * 1e-6):.3e} column-steps/s  "
          f"{gbs:7.1f} GB/s algorithmic  frac {gbs / peak:.3f}")
    for s in ss:
        s.close()
