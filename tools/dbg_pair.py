import sys, os, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import climaland_b200
from climaland_b200 import workloads
from helpers import cuda_solver
for ncol in (1000, 4096, 20000, 61206):
    for oop in (False, True):
        w = workloads.make_workload("energy_hydrology", ncol, N=15, seed=0, topmodel=True)
        ref = cuda_solver(w, kernel_variant=3, layout=2, out_of_place=oop)
        ref.implicit_step(900.0, 3)
        a = ref.get("u_theta_l" if oop else "y_theta_l")
        s = cuda_solver(w, kernel_variant=4, layout=1, out_of_place=oop)
        for rep in range(3):
            st = s.implicit_step(900.0, 3, want_stats=True)
            b = s.get("u_theta_l" if oop else "y_theta_l")
            if oop or rep == 0:
                bad = ~np.isfinite(b)
                err = np.nanmax(np.abs(a - b) / np.abs(a))
                cols = np.unique(np.nonzero(bad)[0])
                print(ncol, oop, rep, "nan_count", st["nan_count"], "nonfinite", bad.sum(), "relerr", err, "bad cols", cols[:10], cols[-5:] if cols.size else "")
        s.close(); ref.close()
