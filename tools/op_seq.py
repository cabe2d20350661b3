#!/usr/bin/env python3
"""One k_step launch per solver operation over a batch (ncu target: per-phase kernels in isolation)."""
import importlib, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hkd-mpc_b200")
wl = importlib.import_module("hkd-mpc_b200.workloads")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = wl.config3(pkg, n)
B = pkg.MultiPhaseDDPBatch(0)
B.set_problems(w.schedules, w.schedule_id)   # launch 0: reset
B.set_initial_condition(w.x0)
B.hybrid_rollout(0.0)      # 1
B.update_nominal()         # 2
B.compute_cost()           # 3
B.lq_approximation()       # 4
B.backward_sweep(0.0)      # 5
B.linear_rollout(1.0)      # 6
B.hybrid_rollout(1.0)      # 7
B.compute_cost()           # 8
print("ok")
