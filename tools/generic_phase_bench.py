#!/usr/bin/env python3
"""Throughput of the generic SinglePhase<xs,us,ys> sweeps on one GPU: backward sweep and linear
rollout of n independent phases, inputs resident in HBM, CUDA-event times of the kernels.  Prints one JSON line per
instantiation with the dense algorithmic FLOP / byte counts per stage and the fractions of the measured peaks."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hkd-mpc_b200")
wl = importlib.import_module("hkd-mpc_b200.workloads")
random_phase = wl.random_phase

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 60
peak = pkg.fp64_peak_tflops(0, 0)
hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6553.6) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6553.6


flop_per_stage, bytes_per_stage = wl.generic_flop_per_stage, wl.generic_bytes_per_stage


for xs, us, ys in [(12, 12, 0), (24, 24, 0), (36, 12, 12)]:
    one = random_phase(xs, us, ys, N, 1, n=8)
    B = pkg.SinglePhaseBatch(xs, us, ys, N, n)
    for nm in B.INPUTS:
        B.set(nm, np.ascontiguousarray(np.broadcast_to(one[nm][None], (n // 8,) + one[nm].shape).reshape((n,) + one[nm].shape[1:])))
    ts, tr = [], []
    for _ in range(5):
        assert B.backward_sweep(1e-3).all(); ts.append(B.last_ms())
        B.linear_rollout(1.0); tr.append(B.last_ms())
    ms, mr = float(np.median(ts[1:])), float(np.median(tr[1:]))
    F, Bt = flop_per_stage(xs, us, ys), bytes_per_stage(xs, us, ys)
    print(json.dumps(dict(instantiation=[xs, us, ys], problems=n, horizon=N, sweep_ms=round(ms, 3), rollout_ms=round(mr, 3),
                          stages_per_s=round(n * N / ms * 1e3), flop_per_stage=F, bytes_per_stage=Bt,
                          sweep_tflops=round(n * N * F / ms / 1e9, 3), fp64_frac=round(n * N * F / ms / 1e9 / peak, 4),
                          sweep_gbps=round(n * N * Bt / ms / 1e6, 1), hbm_frac=round(n * N * Bt / ms / 1e6 / hbm, 4),
                          fp64_peak_tflops=round(peak, 2), hbm_peak_gbps=hbm)))
