"""Multi-GPU plumbing: one process per GPU, the active-column list is cut into contiguous
blocks, no halo and no data-path collective -- columns are independent
(rre.jl:106, energy_hydrology.jl:258, utils.jl:183-188).  torch.distributed is used only to
hand the NCCL unique id to every rank; the reductions themselves (Newton norm, NaN count,
water / energy balance) run inside the library on its own NCCL communicator."""
import numpy as np

from . import workloads as _wl


def shard_range(n_columns, world_size, rank):
    """Contiguous block [lo, hi) of rank `rank`; sizes differ by at most one column."""
    lo = (n_columns * rank) // world_size
    hi = (n_columns * (rank + 1)) // world_size
    return lo, hi


def shard_workload(w, world_size, rank):
    """Slice every per-cell / per-column array of a workload dict to this rank's columns."""
    lo, hi = shard_range(w["ncol"], world_size, rank)
    out = {}
    for k, v in w.items():
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == w["ncol"] and k not in ("z_f", "z_c"):
            out[k] = np.ascontiguousarray(v[lo:hi])
        else:
            out[k] = v
    out["ncol"] = hi - lo
    return out


def attach_communicator(solver, dist=None):
    """Give `solver` the library's NCCL communicator over the default torch.distributed group."""
    if dist is None:
        import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = [type(solver).comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    solver.comm_init(uid[0], world, rank)


def global_norm(local_sumsq, dist=None):
    """sqrt of the all-reduced sum of squares (the ConvergenceChecker norm over all ranks)."""
    import torch
    if dist is None:
        import torch.distributed as dist
    t = torch.tensor([float(local_sumsq)], dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.sqrt().item())


__all__ = ["shard_range", "shard_workload", "attach_communicator", "global_norm"]
_ = _wl
