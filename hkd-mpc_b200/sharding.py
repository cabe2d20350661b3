"""Multi-GPU plumbing: problems are independent, so they are sharded by index across
ranks with NO data-path collective.  torch.distributed is used only
for the launch barrier, the max-over-ranks timing and the gather of convergence
statistics (a few integers/doubles per rank).  Works with the gloo backend on CPU
(tests) and nccl on GPU (bench.py)."""
import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous index range [lo, hi) of rank `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


STAT_FIELDS = ("n_problems", "n_converged", "n_stalled", "n_max_iter", "n_failed", "sum_iters", "max_iters",
               "sum_sweeps", "sum_trials", "sum_sweep_stages")


def local_stats(info, sweep_stages):
    """Convergence statistics of one rank from the hsddp_info records."""
    st = info["status"]
    return np.array([len(st), int((st == 0).sum()), int((st == 1).sum()), int((st == 2).sum()), int((st == 3).sum()),
                     int(info["n_iter"].sum()), int(info["n_iter"].max()) if len(st) else 0, int(info["n_sweeps"].sum()),
                     int(info["n_trials"].sum()), int(sweep_stages)], np.float64)


def gather_stats(stats, device=None):
    """All-gather the per-rank statistics; returns [world, len(STAT_FIELDS)] on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(stats, np.float64)[None]
    t = torch.tensor(np.asarray(stats, np.float64), device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.stack([o.cpu().numpy() for o in out])


def reduce_stats(gathered):
    g = np.asarray(gathered)
    tot = dict(zip(STAT_FIELDS, g.sum(axis=0)))
    tot["max_iters"] = float(g[:, STAT_FIELDS.index("max_iters")].max())
    return tot


def max_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
