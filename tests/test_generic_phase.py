"""the reference's other instantiations SinglePhase<double,12,12,0> and <36,12,12> (ys > 0 output path).

The reference ships no model, cost or problem for them, so the boundary is the phase's storage after LQ_approximation
(plug-in outputs in, sweep outputs out).  CPU part: the run-time-sized oracle restatement (oracle/single_phase_generic.hpp)
is (a) bit-identical to the fixed-size HKD oracle at <24,24,0>, (b) equal to an independent textbook NumPy restatement.
GPU part: the CUDA sweeps (csrc/hsddp_generic.cu, through the C ABI) against the oracle, 1e-9 per row.
"""
import ctypes as C
import numpy as np
import pytest
from conftest import GAIT_PATH, load_pkg, load_workloads

INSTANTIATIONS = [(12, 12, 0), (36, 12, 12), (24, 24, 0)]  # HSDDPSolver/source/SinglePhase.cpp:538-540
TOL = 1e-9


random_phase = load_workloads().random_phase  # synthetic plug-in outputs (hkd-mpc_b200/workloads.py)


def numpy_backward_sweep(xs, us, ys, N, d, reg, Gp, Hp):
    """Independent restatement: textbook formulas, NumPy products, np.linalg.inv, eigenvalue PD test."""
    G = np.zeros((N + 1, xs)); H = np.zeros((N + 1, xs, xs)); K = np.zeros((N, us, xs)); dU = np.zeros((N, us))
    G[N] = d["Phix"] + Gp; H[N] = d["Phixx"] + Hp
    dV1 = dV2 = 0.0
    ok = True
    for k in range(N - 1, -1, -1):
        A, B = d["A"][k], d["B"][k]
        Gn = G[k + 1] + H[k + 1] @ d["Defect"][k + 1]
        Qx = d["lx"][k] + A.T @ Gn
        Qu = d["lu"][k] + B.T @ Gn
        Qxx = d["lxx"][k] + A.T @ H[k + 1] @ A
        Quu = d["luu"][k] + B.T @ H[k + 1] @ B
        Qux = d["lux"][k] + B.T @ H[k + 1] @ A
        if ys:
            Cm, Dm = d["C"][k], d["D"][k]
            Qx = Qx + Cm.T @ d["ly"][k]; Qu = Qu + Dm.T @ d["ly"][k]
            Qxx = Qxx + Cm.T @ d["lyy"][k] @ Cm; Quu = Quu + Dm.T @ d["lyy"][k] @ Dm; Qux = Qux + Dm.T @ d["lyy"][k] @ Cm
        Qxx = Qxx + reg * np.eye(xs); Quu = Quu + reg * np.eye(us)
        if np.linalg.eigvalsh((Quu + Quu.T) / 2 - 1e-9 * np.eye(us)).min() < 0:
            ok = False
            break
        Qi = np.linalg.inv(Quu); Qi = (Qi + Qi.T) / 2
        Qxx = (Qxx + Qxx.T) / 2
        dU[k] = -Qi @ Qu; K[k] = -Qi @ Qux
        G[k] = Qx - Qux.T @ Qi @ Qu; H[k] = Qxx - Qux.T @ Qi @ Qux
        dVk = -Qu @ dU[k]
        dV1 -= dVk; dV2 += dVk
    G[0] = G[0] + H[0] @ d["Defect"][0]
    return dict(success=ok, dU=dU, K=K, G=G, H=H, dV_1=dV1, dV_2=dV2)


def numpy_linear_rollout(xs, us, N, d, eps, dU, K, dx0):
    dX = np.zeros((N + 1, xs)); dV1 = dV2 = 0.0
    dX[0] = dx0 + eps * d["Defect"][0]
    for k in range(N):
        du = eps * dU[k] + K[k] @ dX[k]
        dX[k + 1] = d["A"][k] @ dX[k] + d["B"][k] @ du + eps * d["Defect"][k + 1]
        dV1 += d["lx"][k] @ dX[k] + d["lu"][k] @ du
        dV2 += dX[k] @ d["lxx"][k] @ dX[k] + du @ d["luu"][k] @ du + du @ d["lux"][k] @ dX[k]
    dV1 += d["Phix"] @ dX[N]; dV2 += dX[N] @ d["Phixx"] @ dX[N]
    return dict(dX=dX, dV_1=dV1, dV_2=dV2)


def rows_close(a, b, tol=TOL):
    """relative error per leading row (stage / node), so small rows are not hidden by large ones"""
    a = np.asarray(a, float); b = np.asarray(b, float)
    a2 = a.reshape(a.shape[0], -1) if a.ndim > 1 else a.reshape(-1, 1)
    b2 = b.reshape(a2.shape)
    err = np.abs(a2 - b2).max(axis=1) / np.maximum(np.abs(b2).max(axis=1), 1e-300)
    return float(err.max()) if err.size else 0.0


# ------------------------------------------------------------------ CPU: the oracle restatement itself
def test_generic_oracle_is_bit_identical_to_the_hkd_oracle_at_24_24_0(orc):
    P = orc.Problem(orc.GaitTable(GAIT_PATH("trot")), 0, 0.6)
    assert P.hybrid_rollout(0.0)
    P.update_nominal(); P.compute_cost(); P.lq_approximation()
    assert P.backward_sweep(0.0)
    P.linear_rollout(1.0)
    hor = [p["horizon"] for p in P.phases]
    s0, n0 = sum(hor[:-1]), sum(hor[:-1]) + len(hor) - 1  # first stage / node of the LAST phase (G' = 0, H' = 0)
    N = hor[-1]
    g = lambda nm, a, b: P.get(nm)[a:b]
    d = dict(A=g("A", s0, s0 + N), B=g("B", s0, s0 + N), lx=g("lx", s0, s0 + N), lu=g("lu", s0, s0 + N), lxx=g("lxx", s0, s0 + N),
             luu=g("luu", s0, s0 + N), lux=g("lux", s0, s0 + N), Phix=P.get("Phix")[-1], Phixx=P.get("Phixx")[-1],
             Defect=g("Defect", n0, n0 + N + 1), C=np.zeros((N, 0, 24)), D=np.zeros((N, 0, 24)), ly=np.zeros((N, 0)), lyy=np.zeros((N, 0, 0)))
    r = orc.generic_backward_sweep(24, 24, 0, N, d, 0.0)
    assert r["success"]
    assert np.array_equal(r["K"], g("K", s0, s0 + N))
    assert np.array_equal(r["dU"], g("dU", s0, s0 + N))
    assert np.array_equal(r["H"], g("H", n0, n0 + N + 1))
    assert np.array_equal(r["G"], g("G", n0, n0 + N + 1))
    # linear rollout of the last phase from the dX the multi-phase rollout handed it
    dX = g("dX", n0, n0 + N + 1)
    dx_init = dX[0] - 1.0 * d["Defect"][0]
    lr = orc.generic_linear_rollout(24, 24, 0, N, d, 1.0, r["dU"], r["K"], dx_init)
    assert np.abs(lr["dX"] - dX).max() <= 1e-15 * max(1.0, np.abs(dX).max())  # dx_init is reconstructed to one rounding


@pytest.mark.parametrize("xs,us,ys", INSTANTIATIONS)
def test_generic_oracle_against_independent_numpy_restatement(orc, xs, us, ys):
    N = 15
    for seed in range(3):
        d = random_phase(xs, us, ys, N, 100 * xs + seed)
        rng = np.random.default_rng(seed)
        Gp = rng.normal(size=xs); M = rng.normal(size=(xs, xs)) * 0.2; Hp = M @ M.T
        r = orc.generic_backward_sweep(xs, us, ys, N, d, 1e-3 * seed, Gp, Hp)
        q = numpy_backward_sweep(xs, us, ys, N, d, 1e-3 * seed, Gp, Hp)
        assert r["success"] and q["success"]
        for nm in ("K", "dU", "G", "H"):
            assert rows_close(r[nm], q[nm]) < TOL, nm
        assert abs(r["dV_1"] - q["dV_1"]) < TOL * abs(q["dV_1"]) and abs(r["dV_2"] - q["dV_2"]) < TOL * abs(q["dV_2"])
        dx0 = 0.1 * rng.normal(size=xs)
        a = orc.generic_linear_rollout(xs, us, ys, N, d, 0.5, r["dU"], r["K"], dx0)
        b = numpy_linear_rollout(xs, us, N, d, 0.5, r["dU"], r["K"], dx0)
        assert rows_close(a["dX"], b["dX"]) < TOL
        assert abs(a["dV_1"] - b["dV_1"]) < TOL * max(1.0, abs(b["dV_1"])) and abs(a["dV_2"] - b["dV_2"]) < TOL * max(1.0, abs(b["dV_2"]))


def test_generic_oracle_failed_stage_semantics(orc):
    """An indefinite Quu breaks the loop: false is returned, the stages below keep what the storage held, and
    G[0] += H[0] Defect[0] still runs (SinglePhase.cpp:342-347,365)."""
    xs, us, ys, N = 12, 12, 0, 8
    d = random_phase(xs, us, ys, N, 7)
    d["luu"][3] = -50.0 * np.eye(us)
    r = orc.generic_backward_sweep(xs, us, ys, N, d, 0.0)
    assert not r["success"]
    assert np.all(r["K"][:4] == 0) and np.all(r["K"][4:] != 0)
    q = numpy_backward_sweep(xs, us, ys, N, d, 0.0, np.zeros(xs), np.zeros((xs, xs)))
    assert not q["success"]
    assert rows_close(r["K"][4:], q["K"][4:]) < TOL


def test_unknown_instantiation_is_refused():
    pkg = load_pkg()
    h = C.c_void_p()
    rc = pkg.lib().hsddp_phase_batch_create(0, 10, 4, 0, 5, 1, C.byref(h))
    assert rc == -3 and not h  # HSDDP_ERR_UNSUPPORTED, before any device is touched
    assert b"SinglePhase.cpp:538-540" in pkg.lib().hsddp_last_error()


# ------------------------------------------------------------------ GPU: CUDA sweeps vs the oracle
@pytest.mark.gpu
@pytest.mark.parametrize("xs,us,ys", INSTANTIATIONS)
def test_gpu_generic_sweeps_match_oracle(orc, xs, us, ys):
    pkg = load_pkg()
    n, N = 48, 20
    d = random_phase(xs, us, ys, N, 31 * xs + ys, n=n)
    rng = np.random.default_rng(5)
    Gp = rng.normal(size=(n, xs)); M = rng.normal(size=(n, xs, xs)) * 0.2; Hp = M @ np.swapaxes(M, 1, 2)
    dx0 = 0.1 * rng.normal(size=(n, xs))
    B = pkg.SinglePhaseBatch(xs, us, ys, N, n)
    for nm in B.INPUTS:
        B.set(nm, d[nm])
    ok = B.backward_sweep(2e-3, Gp, Hp)
    assert ok.all()
    got = {nm: B.get(nm) for nm in ("dU", "K", "G", "H", "dV")}
    B.linear_rollout(0.5, dx0)
    dX, dVl = B.get("dX"), B.get("dV")
    worst = 0.0
    for i in range(n):
        di = {k: v[i] for k, v in d.items()}
        r = orc.generic_backward_sweep(xs, us, ys, N, di, 2e-3, Gp[i], Hp[i])
        assert r["success"]
        for nm in ("K", "dU", "G", "H"):
            e = rows_close(got[nm][i], r[nm]); worst = max(worst, e)
            assert e < TOL, (nm, i, e)
        assert abs(got["dV"][i, 0] - r["dV_1"]) < TOL * abs(r["dV_1"]) and abs(got["dV"][i, 1] - r["dV_2"]) < TOL * abs(r["dV_2"])
        a = orc.generic_linear_rollout(xs, us, ys, N, di, 0.5, r["dU"], r["K"], dx0[i])
        e = rows_close(dX[i], a["dX"]); worst = max(worst, e)
        assert e < TOL, ("dX", i, e)
        assert abs(dVl[i, 0] - a["dV_1"]) < TOL * max(1.0, abs(a["dV_1"])) and abs(dVl[i, 1] - a["dV_2"]) < TOL * max(1.0, abs(a["dV_2"]))
    print(f"<{xs},{us},{ys}>: worst relative row error {worst:.2e}")


@pytest.mark.gpu
def test_gpu_generic_failed_stage_matches_oracle(orc):
    pkg = load_pkg()
    xs, us, ys, N, n = 36, 12, 12, 10, 6
    d = random_phase(xs, us, ys, N, 11, n=n)
    d["luu"][1, 4] = -80.0 * np.eye(us)   # problem 1 fails at stage 4
    d["luu"][3, 0] = -80.0 * np.eye(us)   # problem 3 fails at the last stage visited
    B = pkg.SinglePhaseBatch(xs, us, ys, N, n)
    for nm in B.INPUTS:
        B.set(nm, d[nm])
    ok = B.backward_sweep(0.0)
    assert list(ok) == [True, False, True, False, True, True]
    K, G = B.get("K"), B.get("G")
    for i in (1, 3):
        r = orc.generic_backward_sweep(xs, us, ys, N, {k: v[i] for k, v in d.items()}, 0.0)
        assert not r["success"]
        k_fail = 4 if i == 1 else 0
        assert np.all(K[i, :k_fail + 1] == 0)
        assert rows_close(K[i, k_fail + 1:], r["K"][k_fail + 1:]) < TOL
        assert rows_close(G[i, k_fail + 1:], r["G"][k_fail + 1:]) < TOL


@pytest.mark.gpu
def test_gpu_generic_24_24_0_on_hkd_stage_data(orc):
    """The dense <24,24,0> sweep on the HKD oracle's own LQ data of a last phase reproduces the HKD oracle's gains
    (the structure-exploiting HKD kernels are tested against the same oracle in test_gpu_parity.py)."""
    pkg = load_pkg()
    P = orc.Problem(orc.GaitTable(GAIT_PATH("bound")), 40, 0.6)
    assert P.hybrid_rollout(0.0)
    P.update_nominal(); P.compute_cost(); P.lq_approximation()
    assert P.backward_sweep(0.0)
    hor = [p["horizon"] for p in P.phases]
    s0, n0, N = sum(hor[:-1]), sum(hor[:-1]) + len(hor) - 1, hor[-1]
    g = lambda nm, a, b: P.get(nm)[a:b]
    B = pkg.SinglePhaseBatch(24, 24, 0, N, 1)
    for nm in ("A", "B", "lx", "lu", "lxx", "luu", "lux"):
        B.set(nm, g(nm, s0, s0 + N)[None])
    B.set("Phix", P.get("Phix")[-1][None]); B.set("Phixx", P.get("Phixx")[-1][None]); B.set("Defect", g("Defect", n0, n0 + N + 1)[None])
    assert B.backward_sweep(0.0).all()
    assert rows_close(B.get("K")[0], g("K", s0, s0 + N)) < TOL
    assert rows_close(B.get("dU")[0], g("dU", s0, s0 + N)) < TOL
    assert rows_close(B.get("H")[0], g("H", n0, n0 + N + 1)) < TOL
    assert rows_close(B.get("G")[0], g("G", n0, n0 + N + 1)) < TOL


@pytest.mark.gpu
def test_gpu_generic_cpp_shim_example_matches_oracle(orc, tmp_path):
    """examples/generic_sweep.cpp: SinglePhaseSweeps<double,36,12,12> (hkd-mpc_b200/host/MultiPhaseDDP.hpp) over the C ABI."""
    import os
    import subprocess
    from conftest import ROOT
    pkg = load_pkg()
    xs, us, ys, N, n = 36, 12, 12, 9, 5
    d = random_phase(xs, us, ys, N, 77, n=n)
    mats = {"A", "B", "C", "D", "lxx", "luu", "lux", "lyy", "Phixx"}
    path = str(tmp_path / "phase.bin")
    with open(path, "wb") as f:
        for nm in pkg.SinglePhaseBatch.INPUTS:  # the C ABI's order and layout: column-major matrices
            a = np.swapaxes(d[nm], -1, -2) if nm in mats else d[nm]
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
    exe = str(tmp_path / "generic_sweep")
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", os.path.join(ROOT, "examples", "generic_sweep.cpp"), "-L" + libdir, "-lhsddp_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe, path, str(n), str(N), "0.001", "0.5"], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(out) == n
    for i, line in enumerate(out):
        v = dict(t.split("=") for t in line.split())
        di = {k: a[i] for k, a in d.items()}
        r = orc.generic_backward_sweep(xs, us, ys, N, di, 0.001)
        lr = orc.generic_linear_rollout(xs, us, ys, N, di, 0.5, r["dU"], r["K"])
        assert int(v["ok"]) == int(r["success"]) == 1
        for got, want in ((v["dV_1"], r["dV_1"]), (v["dV_2"], r["dV_2"]), (v["K00"], r["K"][0, 0, 0]), (v["lr_dV_1"], lr["dV_1"]), (v["dXN0"], lr["dX"][N, 0])):
            assert abs(float(got) - want) <= TOL * max(abs(want), 1e-3), (i, line)
