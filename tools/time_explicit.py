#!/usr/bin/env python
"""Times the explicit-stage kernels (update_aux!, PhaseChange) on the bench workload's shape:
python tools/time_explicit.py  -> us per call, algorithmic GB/s, fraction of the measured HBM bandwidth."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import climaland_b200 as cl  # noqa: F401
from climaland_b200 import workloads
from helpers import cuda_solver

NCOL, N = 61206, 15
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
w = workloads.make_workload("energy_hydrology", NCOL, N=N, seed=0, topmodel=True)
xp = workloads.make_explicit_params(w, 0)
# algorithmic doubles per cell: update_aux! reads 3 state + 8 parameters + 6 explicit parameters, writes 6;
# the totals kernel re-reads the 3 state fields; PhaseChange reads 3 state + 8 parameters + 3 cache, and
# reads + writes the two tendencies; fused = update_aux! + the 4 tendency accesses
BYTES = {"update_aux": 8 * (17 + 6 + 3), "phase_change_source": 8 * (14 + 4), "update_aux_and_phase_change": 8 * (17 + 6 + 3 + 4)}
for mm, mname, kern in ((0, "fast (warp-uniform)", 0), (0, "fast (per-cell cases)", 1), (1, "libm", 0)):
    ss = []
    for r in range(4):
        s = cuda_solver(w, math_mode=mm)
        s.set_option("explicit_kernel", kern)
        for k, v in xp.items():
            s.set(k, v)
        s.set_explicit_params(**workloads.EXPLICIT_SCALARS)
        ss.append(s)
    for name in ("update_aux", "phase_change_source", "update_aux_and_phase_change"):
        for s in ss:
            getattr(s, name)()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(400):
            getattr(ss[k % 4], name)()
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / 400
        gbs = NCOL * N * BYTES[name] / (us * 1e-6) / 1e9
        print(f"{mname:5s} {name:30s} {us:8.1f} us  {gbs:7.1f} GB/s algorithmic  frac {gbs / peak:.3f}")
    for s in ss:
        s.close()
