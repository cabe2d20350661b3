#!/usr/bin/env python3
"""SURVEY.md §8(d) configurations on one GPU: config 2 (1,024 perturbed trot), config 3 (16,384 mixed gaits),
config 4 (4,096 long-flight bound+jump) and the config-5 sweep (batch 1 .. 131,072, horizon 0.25 .. 1.0 s).
Every line: cold solve of the whole batch, kernel time from CUDA events on the handle's stream (inputs resident)."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hkd-mpc_b200")
wl = importlib.import_module("hkd-mpc_b200.workloads")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"


def run(name, w):
    t0 = time.time()
    B = pkg.MultiPhaseDDPBatch(0)
    B.set_problems(w.schedules, w.schedule_id)
    B.set_initial_condition(w.x0)
    ms = []
    for rep in range(3 if w.n <= 16384 else 2):
        B.reset(); B.solve()
        ms.append(B.last_solve_ms())
    info = B.info()
    st = np.bincount(info["status"], minlength=4)
    rec = dict(config=name, problems=w.n, plan_s=w.plan, schedules=len(w.schedules), stages_mean=float(np.mean([w.schedules[s].n_stages for s in w.schedule_id])),
               phases_mean=float(np.mean([w.schedules[s].n_phases for s in w.schedule_id])), kernel_ms=float(min(ms[1:])), solves_per_s=w.n / (min(ms[1:]) * 1e-3),
               iters_mean=float(info["n_iter"].mean()), iters_max=int(info["n_iter"].max()), sweeps=int(info["n_sweeps"].sum()), trials=int(info["n_trials"].sum()),
               converged=int(st[0]), stalled=int(st[1]), max_iter=int(st[2]), reg_overflow=int(st[3]), build_s=round(time.time() - t0, 1))
    print(json.dumps(rec), flush=True)
    del B


run("config1", wl.config1(pkg))
run("config2", wl.config2(pkg, 1024))
run("config4", wl.config4(pkg, 4096))
run("config3", wl.config3(pkg, 16384))
for plan in (0.25, 0.5, 0.75, 1.0):
    run("config5-horizon", wl.config3(pkg, 4096, plan))
for n in ((8, 64, 512) if quick else (8, 64, 512, 65536, 131072)):
    run("config5-batch", wl.config3(pkg, n, 0.6))
