#!/usr/bin/env python
"""Where a warp of the bench kernel spends its cycles: a tuning build with -DCLB_PHASE_CLOCKS
(tools/build_variant.sh pc -DCLB_PHASE_CLOCKS; CLB_LIBRARY_PATH=build/exp/pc.so) adds up the SM clock ticks lane 0 of
every warp spends in each phase of a tile; this prints them per tile and as shares.  The clock reads perturb the
schedule (they are barriers for the instruction scheduler), so the launch time is printed beside the normal one."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import climaland_b200 as cl  # noqa: F401
from climaland_b200 import workloads, _lib
from helpers import cuda_solver
NAMES = ["tile prologue", "wait for the tile", "set-up", "closures + T", "neighbour exchange", "faces, residuals, rows",
         "W11 solve", "update of theta", "W21 x1 + W22 solve", "next tile's request", "stores", "  (request: syncwarp + proxy fence)",
         "set-up: lagged loads, 1/rho_c", "set-up: exchange", "set-up: constants", "set-up: W22 rows + factorisation"]
L = _lib.lib()
fn = L.clb_debug_phase_clocks
fn.argtypes = [C.POINTER(C.c_ulonglong)]
for model, iters, dt in (("energy_hydrology", 3, 900.0), ("richards", 2, 1800.0)):
    ncol = int(os.environ.get('PC_NCOL', 61206))
    w = workloads.make_workload(model, ncol, N=15, seed=1, topmodel=True)
    ss = [cuda_solver(w, out_of_place=True) for _ in range(4)]
    for s in ss: s.implicit_step(dt, iters)
    out = (C.c_ulonglong * 16)()
    fn(out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    e0.record()
    for k in range(reps): ss[k % 4].implicit_step(dt, iters)
    e1.record(); torch.cuda.synchronize()
    fn(out)
    us = 1e3 * e0.elapsed_time(e1) / reps
    tiles = (ncol + 7) // 8 * reps
    tot = sum(out)
    print(f"{model}: {us:.1f} us per launch (instrumented); {tot / tiles:.0f} ticks per tile, {tot / reps / int(os.environ.get('PC_WARPS', 1184)):.0f} per warp and launch")
    for n, v in zip(NAMES, out):
        if v: print(f"  {n:26s} {v / tiles:9.0f} ticks/tile  {100.0 * v / tot:5.1f} %")
    for s in ss: s.close()
