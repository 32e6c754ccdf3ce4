#!/usr/bin/env python
"""Column sweep of the fused implicit stage (BASELINE.json configs[2]: Richards synthetic column
sweep 1e4-1e7 columns x 15/50 layers; plus the two ~1 degree EnergyHydrology column counts).

  python tools/sweep.py [--max-cols 1e7] [--steps 200] > profiles/rN/sweep.txt

One line per case: kernel the library chose, us per stage, column-steps/s, algorithmic GB/s and
its fraction of the measured HBM copy bandwidth (MEASURED_PEAKS.json).  Inputs: a base block of
<= 1e5 synthetic columns (workloads.make_workload) tiled up to the case's column count, resident in
HBM; steps rotate over enough independent handles to exceed the 126 MB L2; CUDA events on the
launching stream.  The library's result on the base block is checked against the CPU oracle
(1e-12) before timing, so a timed configuration is a parity-checked one.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

BASE = 100_000
VARIANTS = {1: "thread/column registers", 2: "thread/column generic", 3: "lane per cell", 4: "lane quad",
            5: "lane quad pipelined", 6: "lane octet pipelined"}


def tiled(a, ncol):
    if a.shape[0] == ncol:
        return a
    reps = -(-ncol // a.shape[0])
    return np.ascontiguousarray(np.concatenate([a] * reps, axis=0)[:ncol])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-cols", type=float, default=1e7)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--cases", default="all")
    args = ap.parse_args()
    import torch
    import climaland_b200 as cl
    from climaland_b200 import workloads
    from helpers import oracle_problem, rel_err

    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    cases = []
    for N in (15, 50):
        for ncol in (10_000, 100_000, 1_000_000, 10_000_000):
            if ncol <= args.max_cols:
                cases.append(("richards", N, ncol, False, 1800.0, 2))
    for ncol in (61_206, 64_800):
        cases.append(("energy_hydrology", 15, ncol, True, 900.0, 3))
    cases.append(("energy_hydrology", 50, 100_000, True, 900.0, 3))
    cases.append(("richards", 15, 61_206, True, 1800.0, 2))
    # the masked variant of SURVEY 8(d): a 64 800-column lat-long domain with 30 % land -- the handle holds the active
    # columns only (mask_test.jl:53-61), so the stage runs on 19 440 compacted columns
    cases.append(("energy_hydrology", 15, 19_440, True, 900.0, 3))
    # BASELINE configs[0]'s boundary conditions (MoistureStateBC top, FreeDrainage bottom): fluxes at the iterate
    cases.append(("richards-bc", 15, 100_000, False, 1800.0, 2))
    stream = torch.cuda.Stream()
    print(f"# peak = {peak} GB/s (measured copy bandwidth); columns tiled from a {BASE}-column block")
    print(f"{'model':17s} {'N':>3s} {'columns':>9s} {'kernel':26s} {'us/stage':>10s} {'col-steps/s':>12s} {'GB/s':>8s} {'frac':>6s}  parity")
    for model, N, ncol, topm, dt, iters in cases:
        nb = min(ncol, BASE)
        live_bc = model == "richards-bc"
        label, model = model, model.split("-")[0]
        w = workloads.make_workload(model, nb, N=N, seed=7, topmodel=topm)
        bc = {}
        if live_bc:
            w["theta_bc_top"] = w["nu"][:, -1] - np.random.default_rng(3).uniform(1e-3, 0.1, nb)
            bc = dict(top_bc=1, bottom_bc=1)
        # parity of the base block against the oracle (test infrastructure: the checker only)
        P, U, p = oracle_problem(w, nthreads=os.cpu_count() or 1, **bc)
        P.implicit_step(U, dt, iters, p=p)
        bytes_cs = workloads.algorithmic_bytes(model, N, topmodel=topm)
        per_handle = ncol * N * (10 if model == "richards" else 17) * 8
        if args.cases != "all" and args.cases not in f"{label}-{N}":
            continue
        replicas = int(max(1, min(16, -(-160e6 // per_handle))))
        solvers = []
        mdl = cl.RICHARDS if model == "richards" else cl.ENERGY_HYDROLOGY
        for r in range(replicas):
            s = cl.SoilColumnSolver(model=mdl, n_columns=ncol, z_f=w["z_f"], z_c=w["z_c"], has_topmodel_source=topm,
                                    stream=stream.cuda_stream, out_of_place=True, **bc)
            for k, v in w.items():
                if k.lower() in cl.FIELDS:
                    s.set(k, tiled(np.asarray(v), ncol))
            solvers.append(s)
        with torch.cuda.stream(stream):
            solvers[0].implicit_step(dt, iters)
            got = solvers[0].get("u_theta_l")[:nb]
            err = rel_err(got, U.theta_l)
            for k in range(max(5, replicas)):
                solvers[k % replicas].implicit_step(dt, iters)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps = max(10, min(args.steps, int(2e9 / (ncol * N * iters)) + 10))
            e0.record(stream)
            for k in range(steps):
                solvers[k % replicas].implicit_step(dt, iters)
            e1.record(stream)
            torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / steps
        cps = ncol / (us * 1e-6)
        gbs = cps * bytes_cs / 1e9
        name = VARIANTS.get(solvers[0].last_variant(), "?")
        print(f"{label:17s} {N:3d} {ncol:9d} {name:26s} {us:10.1f} {cps:12.4g} {gbs:8.1f} {gbs / peak:6.3f}  {err:.1e}",
              flush=True)
        assert err <= 1e-10, f"parity {err}"  # a whole Newton stage; per-call parity (1e-12) is what tests/ pins
        for s in solvers:
            s.close()
        del solvers
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
