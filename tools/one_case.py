#!/usr/bin/env python
"""A few fused stages of one (model, N, columns) case -- the target of an ncu capture:
ncu --set full -k regex:k_step_lanes -s 6 -c 1 python tools/one_case.py richards 15 100000"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import climaland_b200 as cl  # noqa: F401
from climaland_b200 import workloads
from helpers import cuda_solver
model, N, ncol = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dt, iters = (1800.0, 2) if model == "richards" else (900.0, 3)
w = workloads.make_workload(model, ncol, N=N, seed=1, topmodel=True)
ss = [cuda_solver(w, out_of_place=True) for _ in range(2)]
for k in range(12):
    ss[k % 2].implicit_step(dt, iters)
torch.cuda.synchronize()
for s in ss:
    s.close()
