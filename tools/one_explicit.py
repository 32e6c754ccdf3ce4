#!/usr/bin/env python
"""A few whole resident soil steps of the bench workload -- the target of an ncu capture of the explicit-stage kernels:
ncu --set full -k regex:k_explicit_cells -s 4 -c 1 python tools/one_explicit.py [explicit_kernel]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
import climaland_b200 as cl
from climaland_b200 import workloads
kern = int(sys.argv[1]) if len(sys.argv) > 1 else 0
w = workloads.make_workload("energy_hydrology", 61206, N=15, seed=1, topmodel=True)
rng = np.random.default_rng(5)
ss = []
for r in range(2):
    s = cl.SoilColumnSolver.from_workload(w)
    s.set_option("explicit_kernel", kern)
    for k, v in workloads.make_explicit_params(w, r).items():
        s.set(k, v)
    s.set_explicit_params(**workloads.EXPLICIT_SCALARS)
    s.set("f_max", rng.uniform(0.2, 0.6, 61206))
    s.set("precip", -rng.uniform(0, 4e-7, 61206))
    s.set_runoff_params(f_over=3.28, R_sb=1.484e-7, depth=50.0)
    ss.append(s)
for k in range(8):
    ss[k % 2].soil_step(900.0, 3)
torch.cuda.synchronize()
for s in ss:
    s.close()
