#!/usr/bin/env python3
"""Developer check on a GPU box: step-level and full-solve parity of the CUDA path
against the CPU oracle, plus a quick throughput probe.  Prints diagnostics; the
assertions live in tests/."""
import importlib
import os
import sys
import time
import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as op  # noqa: E402

pkg = importlib.import_module("hkd-mpc_b200")
GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def flat_states(P, arr):  # [nodes,24] of the oracle == first n_nodes rows of the GPU layout
    return arr


def step_parity(gait="trot", k0=0, plan=0.6):
    T = op.GaitTable(os.path.join(GOLD, f"gait_{gait}.npz"))
    R = pkg.QuadReference(os.path.join(GOLD, f"gait_{gait}.npz"))
    P = op.Problem(T, k0, plan)
    S = pkg.Schedule(R, k0, plan)
    B = pkg.MultiPhaseDDPBatch(0)
    B.set_problems([S], [0])
    B.set_initial_condition(P.x0[None])
    N, NN = S.n_stages, S.n_nodes
    print(f"== step parity {gait} k0={k0} plan={plan}: phases {S.horizon}")
    # initial rollout
    okc = P.hybrid_rollout(0.0); okg = B.hybrid_rollout(0.0)
    print("rollout0 ok", okc, okg[0], "X", rel(B.get("X")[0, :NN], P.get("X")), "Defect", rel(B.get("Defect")[0, :NN], P.get("Defect")))
    P.update_nominal(); B.update_nominal()
    for it in range(3):
        P.compute_cost(); B.compute_cost()
        sc = P.scalars(); sg = B.scalars()[0]
        print(f"it{it} cost", sc["actual_cost"], sg[0], "rel", abs(sc["actual_cost"] - sg[0]) / abs(sc["actual_cost"]), "feas", sc["feas"], sg[2])
        P.lq_approximation(); B.lq_approximation()
        for nm in ("A", "B", "lx", "lu", "lxx", "luu"):
            print("   ", nm, rel(B.get(nm)[0, :N], P.get(nm)))
        okc = P.backward_sweep(0.0); okg = B.backward_sweep(0.0)
        print("   sweep ok", okc, okg[0], "K", rel(B.get("K")[0, :N], P.get("K")), "dU", rel(B.get("dU")[0, :N], P.get("dU")),
              "G0", rel(B.get("G0")[0], P.get("G")[0]), "H0", rel(B.get("H0")[0], P.get("H")[0]))
        sc = P.scalars(); sg = B.scalars()[0]
        print("   sweep dV", sc["dV_1"], sg[3], sc["dV_2"], sg[4])
        P.linear_rollout(1.0); B.linear_rollout(1.0)
        sc = P.scalars(); sg = B.scalars()[0]
        print("   linear dX", rel(B.get("dX")[0, :NN], P.get("dX")), "dV_1", sc["dV_1"], sg[3], "dV_2", sc["dV_2"], sg[4])
        eps = 1.0 if it == 0 else 0.1
        okc = P.hybrid_rollout(eps); okg = B.hybrid_rollout(eps)
        print("   rollout", eps, okc, okg[0], "X", rel(B.get("X")[0, :NN], P.get("X")), "U", rel(B.get("U")[0, :N], P.get("U")),
              "Defect", rel(B.get("Defect")[0, :NN], P.get("Defect")), "g", rel(B.get("g")[0, :N], P.get("g")) if np.abs(P.get("g")).max() > 0 else 0)
        sc = P.scalars(); sg = B.scalars()[0]
        print("   maxt", sc["max_tconstr"], sg[5], "maxp", sc["max_pconstr"], sg[6])
        P.update_nominal(); B.update_nominal()


def solve_parity(gait="trot", k0=0, plan=0.6, n_pert=4, seed=0):
    T = op.GaitTable(os.path.join(GOLD, f"gait_{gait}.npz"))
    R = pkg.QuadReference(os.path.join(GOLD, f"gait_{gait}.npz"))
    S = pkg.Schedule(R, k0, plan)
    rng = np.random.default_rng(seed)
    x0s = []
    base = S.default_x0()
    for i in range(n_pert):
        x = base.copy()
        if i > 0:
            x[0:3] += rng.uniform(-0.05, 0.05, 3); x[3:6] += rng.uniform(-0.02, 0.02, 3)
            x[6:9] += rng.uniform(-0.2, 0.2, 3); x[9:12] += rng.uniform(-0.1, 0.1, 3)
            x[12:] = pkg.compute_hkd_state(x[0:3], x[3:6], np.array([0, -0.8, 1.6] * 4), S.contact[0])
        x0s.append(x)
    x0s = np.array(x0s)
    B = pkg.MultiPhaseDDPBatch(0)
    B.set_problems([S], np.zeros(n_pert, np.int32))
    B.set_initial_condition(x0s)
    t0 = time.time(); B.solve(); t1 = time.time()
    info = B.info(); tr = B.trace()
    Xb, Ub, K, dU = B.get("Xbar"), B.get("Ubar"), B.get("K"), B.get("dU")
    print(f"== solve parity {gait} k0={k0} plan={plan}: gpu solve wall {t1 - t0:.3f}s kernel {B.last_solve_ms():.2f} ms")
    for i in range(n_pert):
        P = op.Problem(T, k0, plan); P.x0 = x0s[i]
        s, otr = P.solve()
        n = int(s["n_iter"])
        same_iter = (n == info["n_iter"][i]) and (int(s["status"]) == info["status"][i])
        same_eps = np.array_equal(otr[:, 9], tr[i, :n, 9])
        cost_rel = np.abs(otr[:, 11] - tr[i, :n, 11]).max() / np.abs(otr[:, 11]).max() if n else 0
        print(f"  p{i}: iters {n}/{info['n_iter'][i]} status {int(s['status'])}/{info['status'][i]} same_eps {same_eps} "
              f"cost {s['cost']:.9f}/{info['cost'][i]:.9f} trace_cost_rel {cost_rel:.2e} Xbar {rel(Xb[i, :S.n_nodes], P.get('Xbar')):.2e} "
              f"Ubar {rel(Ub[i, :S.n_stages], P.get('Ubar')):.2e} K {rel(K[i, :S.n_stages], P.get('K')):.2e} dU {rel(dU[i, :S.n_stages], P.get('dU')):.2e}")
        if not (same_iter and same_eps):
            print("   oracle eps", otr[:, 9]); print("   gpu eps   ", tr[i, :info['n_iter'][i], 9])
            print("   oracle cost_after", otr[:, 11]); print("   gpu cost_after   ", tr[i, :info['n_iter'][i], 11])


def throughput(n=1024, gait="trot"):
    R = pkg.QuadReference(os.path.join(GOLD, f"gait_{gait}.npz"))
    S = pkg.Schedule(R, 0, 0.6)
    rng = np.random.default_rng(1)
    x0 = np.tile(S.default_x0(), (n, 1))
    x0[:, 0:3] += rng.uniform(-0.05, 0.05, (n, 3)); x0[:, 3:6] += rng.uniform(-0.02, 0.02, (n, 3))
    x0[:, 6:9] += rng.uniform(-0.2, 0.2, (n, 3)); x0[:, 9:12] += rng.uniform(-0.1, 0.1, (n, 3))
    for i in range(n):
        x0[i, 12:] = pkg.compute_hkd_state(x0[i, 0:3], x0[i, 3:6], np.array([0, -0.8, 1.6] * 4), S.contact[0])
    B = pkg.MultiPhaseDDPBatch(0)
    B.set_problems([S], np.zeros(n, np.int32))
    B.set_initial_condition(x0)
    for rep in range(3):
        B.reset(); B.solve()
        ms = B.last_solve_ms()
        info = B.info()
        print(f"== throughput n={n}: {ms:.2f} ms -> {n / ms * 1e3:.0f} solves/s; iters mean {info['n_iter'].mean():.2f} "
              f"sweeps {info['n_sweeps'].sum()} status hist {np.bincount(info['status'], minlength=4)}")


if __name__ == "__main__":
    print("fp64 DFMA peak TFLOP/s", pkg.fp64_peak_tflops(0, 0), "DMMA", pkg.fp64_peak_tflops(0, 1))
    step_parity("trot", 0, 0.6)
    solve_parity("trot", 0, 0.6, 4)
    solve_parity("bound", 250, 0.6, 2)
    solve_parity("pronk", 100, 0.6, 2)
    throughput(1024)
