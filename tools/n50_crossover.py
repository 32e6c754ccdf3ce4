import os, sys
sys.path[:0] = [".", "oracle", "tests"]
import numpy as np, torch
import climaland_b200 as cl
from climaland_b200 import workloads
w = workloads.make_workload("richards", 100000, N=50, seed=7)
def tiled(a, n):
    reps = -(-n // a.shape[0]); return np.ascontiguousarray(np.concatenate([a]*reps, axis=0)[:n])
for ncol in (150000, 250000, 400000, 600000):
    for kv in (6, 2):
        s = cl.SoilColumnSolver(model=cl.RICHARDS, n_columns=ncol, z_f=w["z_f"], z_c=w["z_c"], out_of_place=True, kernel_variant=kv)
        for k, v in w.items():
            if k.lower() in cl.FIELDS: s.set(k, tiled(np.asarray(v), ncol))
        for _ in range(3): s.implicit_step(1800.0, 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): s.implicit_step(1800.0, 2)
        e1.record(); torch.cuda.synchronize()
        us = 1e3*e0.elapsed_time(e1)/20
        print(ncol, "octet" if kv == 6 else "generic", "%.1f us" % us, "%.3g col-steps/s" % (ncol/(us*1e-6)), flush=True)
        s.close()
