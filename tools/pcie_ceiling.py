#!/usr/bin/env python
"""tools/pcie_ceiling.py: what the box's host side can deliver to N GPUs at once -- the ceiling of the bench's e2e leg.

Plain cudaMemcpyAsync (torch copy_ on two streams) between pinned host buffers and device buffers of the sizes the
e2e routes move per step: H2D alone, D2H alone, both directions at once; every rank runs the same loop at the same time
(barrier before each leg), so the per-rank rates under contention and their sum are what N concurrent host-buffer
steps can at best get.  Run alone or under torchrun:
    python tools/pcie_ceiling.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_ceiling.py
Prints one JSON line per leg (rank 0)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def leg(name, h2d_mb, d2h_mb, reps=20):
    n_in, n_out = int(h2d_mb * 1e6 / 8), int(d2h_mb * 1e6 / 8)
    hin = torch.empty(max(n_in, 1), dtype=torch.float64, pin_memory=True).fill_(1.0)
    hout = torch.empty(max(n_out, 1), dtype=torch.float64, pin_memory=True)
    din = torch.empty(max(n_in, 1), dtype=torch.float64, device="cuda")
    dout = torch.ones(max(n_out, 1), dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        if n_in:
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
        if n_out:
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)

    for _ in range(3):
        once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / reps
    ts = torch.tensor([t], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    tmax = float(ts.item())
    if rank == 0:
        print(json.dumps({"leg": name, "n_gpus": world, "h2d_mb": h2d_mb, "d2h_mb": d2h_mb, "ms_max_over_ranks": 1e3 * tmax,
                          "per_rank_gbs": (h2d_mb + d2h_mb) / 1e3 / tmax, "aggregate_gbs": world * (h2d_mb + d2h_mb) / 1e3 / tmax}),
              flush=True)


# the stage-only route (clb_implicit_step_host) and the whole-step route (clb_soil_step_host) of the ~1 degree workload
leg("h2d_only_55.8MB", 55.8, 0.0)
leg("d2h_only_15.7MB", 0.0, 15.7)
leg("stage_route_both_55.8+15.7MB", 55.8, 15.7)
leg("h2d_only_23.5MB", 23.5, 0.0)
leg("d2h_only_23.0MB", 0.0, 23.0)
leg("whole_step_route_both_23.5+23.0MB", 23.5, 23.0)
if world > 1:
    dist.destroy_process_group()
