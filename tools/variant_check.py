#!/usr/bin/env python
"""tools/variant_check.py [variant]: parity against the oracle (2 003 columns, one stage) and the launch time of the
bench workload for the library CLB_LIBRARY_PATH points at (tuning aid for build/exp/*.so; variant 5 = pipelined quad)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import climaland_b200 as cl  # noqa: F401
from climaland_b200 import workloads
from helpers import cuda_solver, oracle_problem
import oracle as orc
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 5
models = os.environ.get("VC_MODELS", "energy_hydrology,richards").split(",")
tag = os.path.basename(os.environ.get("CLB_LIBRARY_PATH", "in-tree"))
for model, iters, dt in (("energy_hydrology", 3, 900.0), ("richards", 2, 1800.0)):
    if model not in models: continue
    w = workloads.make_workload(model, 2003, N=15, seed=3, topmodel=True)
    P, U, p = oracle_problem(w, nthreads=8)
    P.implicit_step(U, dt, iters, p=p)
    s = cuda_solver(w, kernel_variant=variant)
    s.implicit_step(dt, iters)
    errs = []
    for name in (("theta_l", "rho_e_int") if model == "energy_hydrology" else ("theta_l",)):
        a, b = s.get("y_" + name), getattr(U, name)
        errs.append(float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6 * np.max(np.abs(b))))))
    s.close()
    ncol = int(os.environ.get("VC_NCOL", 61206))
    w = workloads.make_workload(model, ncol, N=15, seed=1, topmodel=True)
    ss = [cuda_solver(w, out_of_place=True, kernel_variant=variant) for _ in range(4)]
    for s in ss: s.implicit_step(dt, iters)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(400): ss[k % 4].implicit_step(dt, iters)
        e1.record(); torch.cuda.synchronize()
        best = min(best, 1e3 * e0.elapsed_time(e1) / 400)
    B = workloads.algorithmic_bytes(model, 15, True)
    print(f"{tag:14s} {model:17s} elementwise err {max(errs):.2e}  {best:7.2f} us  frac {ncol * B / (best * 1e-6) / 6545.9e9:.3f}", flush=True)
    for s in ss: s.close()
