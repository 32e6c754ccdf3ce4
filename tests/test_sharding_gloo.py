"""world_size-2 `gloo` test of the multi-GPU host logic on CPU: sharding by contiguous column
blocks needs no halo (a sharded run equals the unsharded run bit for bit) and the only
collectives are the scalar reductions (Newton norm over all ranks, balance sums).  The CPU
oracle stands in for the per-rank compute here -- the CUDA path cannot run without a GPU --
so this covers sharding, reduction semantics and rank-0 reporting, not the kernels."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import climaland_b200  # noqa: F401
    from climaland_b200 import parallel, workloads
    from helpers import oracle_problem
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = workloads.make_workload("energy_hydrology", 101, N=15, seed=9, topmodel=True)
    ws = parallel.shard_workload(w, world, rank)
    P, U, p = oracle_problem(ws)
    it, nrm_local = P.implicit_step(U, 900.0, 3, p=p)
    nrm = parallel.global_norm(nrm_local ** 2, dist)
    dz = np.diff(w["z_f"])
    t = torch.tensor([float((U.theta_l @ dz).sum()), float(U.intF_w.sum())], dtype=torch.float64)
    dist.all_reduce(t)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), theta=U.theta_l, rho_e=U.rho_e_int, nrm=nrm, bal=t.numpy(),
             lo_hi=np.array(parallel.shard_range(w["ncol"], world, rank)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_step_equals_single_rank(tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import climaland_b200  # noqa: F401
    from climaland_b200 import workloads
    from helpers import oracle_problem
    w = workloads.make_workload("energy_hydrology", 101, N=15, seed=9, topmodel=True)
    P, U, p = oracle_problem(w)
    it, nrm = P.implicit_step(U, 900.0, 3, p=p)
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    assert parts[0]["lo_hi"][1] == parts[1]["lo_hi"][0]
    theta = np.concatenate([q["theta"] for q in parts])
    rho_e = np.concatenate([q["rho_e"] for q in parts])
    assert np.array_equal(theta, U.theta_l) and np.array_equal(rho_e, U.rho_e_int)  # no halo: bit-identical
    for q in parts:  # every rank holds the same global reductions
        assert abs(q["nrm"] - nrm) <= 1e-12 * nrm
        dz = np.diff(w["z_f"])
        assert abs(q["bal"][0] - (U.theta_l @ dz).sum()) <= 1e-12 * abs(q["bal"][0])
