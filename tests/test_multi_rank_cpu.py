"""N > 1 host-side logic on CPU: world_size-2 gloo run of the index sharding and the
convergence-statistics gather that bench.py uses (no data-path collective exists)."""
import importlib
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sh = importlib.import_module("hkd-mpc_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sh.shard_range(101, rank, world)
    info = np.zeros(hi - lo, dtype=[("status", "i4"), ("n_iter", "i4"), ("n_sweeps", "i4"), ("n_trials", "i4")])
    idx = np.arange(lo, hi)
    info["status"] = idx % 3
    info["n_iter"] = 5 + idx % 7
    info["n_sweeps"] = info["n_iter"]
    info["n_trials"] = 2 * info["n_iter"]
    g = sh.gather_stats(sh.local_stats(info, 60 * int(info["n_sweeps"].sum())))
    tmax = sh.max_over_ranks(10.0 + rank)
    dist.barrier()
    q.put((rank, lo, hi, g, tmax))
    dist.destroy_process_group()


def test_shard_ranges_partition_the_batch():
    sys.path.insert(0, ROOT)
    sh = importlib.import_module("hkd-mpc_b200.sharding")
    for n in (1, 7, 16384, 131072):
        for world in (1, 2, 4, 8):
            r = [sh.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_gather():
    sys.path.insert(0, ROOT)
    sh = importlib.import_module("hkd-mpc_b200.sharding")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, lo0, hi0, g0, t0), (r1, lo1, hi1, g1, t1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 51, 51, 101)
    assert np.array_equal(g0, g1) and g0.shape == (2, len(sh.STAT_FIELDS))
    tot = sh.reduce_stats(g0)
    idx = np.arange(101)
    assert tot["n_problems"] == 101 and tot["n_converged"] == (idx % 3 == 0).sum()
    assert tot["sum_iters"] == (5 + idx % 7).sum() and tot["max_iters"] == 11
    assert t0 == t1 == 11.0
