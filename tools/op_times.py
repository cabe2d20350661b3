#!/usr/bin/env python3
"""Per-operation kernel times with ALL blocks in the same phase (k_step launches), to compare against the
phase-mixed persistent k_solve: the instruction-fetch experiment of DESIGN.md §5."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hkd-mpc_b200")
wl = importlib.import_module("hkd-mpc_b200.workloads")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = sys.argv[2] if len(sys.argv) > 2 else "config3"
w = getattr(wl, cfg)(pkg, n)
B = pkg.MultiPhaseDDPBatch(0)
B.set_problems(w.schedules, w.schedule_id)
B.set_initial_condition(w.x0)
B.reset(); B.solve(); info = B.info()
print(f"{cfg} n={n}: k_solve {B.last_solve_ms():.2f} ms; iters {info['n_iter'].sum()} sweeps {info['n_sweeps'].sum()} trials {info['n_trials'].sum()}")
B.reset()
B.hybrid_rollout(0.0); B.update_nominal(); B.compute_cost(); B.lq_approximation(); B.backward_sweep(0.0); B.linear_rollout(1.0)

def timeit(name, fn, reps=5):
    fn(); B.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    B.sync()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    print(f"  {name:18s} {ms:8.3f} ms per launch  ({ms * 1e3 / n:7.3f} us per problem)")
    return ms
t = {}
t["cost"] = timeit("compute_cost", lambda: B.compute_cost())
t["lq"] = timeit("lq_approximation", lambda: B.lq_approximation())
t["sweep"] = timeit("backward_sweep", lambda: B.backward_sweep(0.0))
t["linear"] = timeit("linear_rollout", lambda: B.linear_rollout(1.0))
t["rollout"] = timeit("hybrid_rollout(.1)", lambda: B.hybrid_rollout(0.1))
t["nominal"] = timeit("update_nominal", lambda: B.update_nominal())
it, sw, tr = float(info['n_iter'].sum()), float(info['n_sweeps'].sum()), float(info['n_trials'].sum())
est = (it * (t["cost"] + t["lq"] + t["linear"]) + sw * t["sweep"] + tr * (t["rollout"] + t["cost"]) + it * t["nominal"]) / n
print(f"  phase-homogeneous estimate of the whole solve: {est:.2f} ms  (includes one launch sync per op)")
