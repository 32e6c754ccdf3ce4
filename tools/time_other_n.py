#!/usr/bin/env python
"""What CLB_VARIANT_AUTO gives for level counts other than the three the lane kernels are instantiated for."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import climaland_b200 as cl  # noqa: F401
from climaland_b200 import workloads
from helpers import cuda_solver
VAR = {1: "thread/column registers", 2: "thread/column generic", 3: "lane per cell", 4: "lane quad", 5: "lane quad pipelined", 6: "lane octet"}
for model, iters, dt in (("richards", 2, 1800.0), ("energy_hydrology", 3, 900.0)):
    # AUTO: lane per cell (N < 15), quad (15, 16), octet (17..64); a trailing "g": the generic kernel forced, for comparison
    cases = [int(a) if a.isdigit() else a for a in sys.argv[1:]] or [10, 15, 16, 20, 25, 30, 40, 50, 56, "56g", 64, "64g"]
    for case in cases:
        N = int(str(case).rstrip("g"))
        ncol = 100_000
        w = workloads.make_workload(model, ncol, N=N, seed=1, topmodel=True)
        ss = [cuda_solver(w, out_of_place=True, kernel_variant=2 if str(case).endswith("g") else 0) for _ in range(2)]
        for s in ss: s.implicit_step(dt, iters)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(40): ss[k % 2].implicit_step(dt, iters)
        e1.record(); torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / 40
        print(f"{model:17s} N={N:3d} {VAR[ss[0].last_variant()]:24s} {us:8.1f} us  {ncol/(us*1e-6):.3e} column-steps/s  "
              f"{ncol*N*iters/(us*1e-6):.3e} cell-iterations/s", flush=True)
        for s in ss: s.close()
