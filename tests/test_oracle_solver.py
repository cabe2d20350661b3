"""Oracle self-checks at the solver level (CPU).

The reference ships no tests or recorded outputs for MultiPhaseDDP::solve
(SURVEY.md §4): parity is UNPINNED at this level.  What can be checked:
  * the Eigen-semantics pieces (pivoted LDLT sign test, partial-pivot LU inverse) against numpy
  * internal invariants of one DDP iteration
  * regression against the committed oracle-generated records (tests/golden/oracle_solves.npz)
  * agreement of the oracle's two model back ends (reference CasADi vs port) over whole solves
"""
import ctypes as C
import os
import numpy as np
import pytest
from conftest import GOLDEN, GAIT_PATH, rel_err


def _table(orc, gait):
    return orc.GaitTable(GAIT_PATH(gait))


def test_ldlt_sign_test_matches_eigenvalues(orc):
    rng = np.random.default_rng(0)
    for t in range(200):
        M = rng.normal(size=(24, 24))
        S = M @ M.T + (rng.uniform(-30, 5) if t % 2 else 1e-3) * np.eye(24)
        lam = np.linalg.eigvalsh(S).min()
        if abs(lam) < 1e-8:
            continue
        Sc = np.asfortranarray(S)
        got = orc.lib().orc_ldlt_is_positive(Sc.ctypes.data_as(C.POINTER(C.c_double)))
        assert bool(got) == (lam > 0)


def test_partial_pivot_inverse(orc):
    rng = np.random.default_rng(1)
    for _ in range(20):
        M = rng.normal(size=(24, 24))
        S = np.asfortranarray(M @ M.T + np.eye(24))
        out = np.zeros((24, 24), order="F")
        orc.lib().orc_inverse(S.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)))
        assert rel_err(out, np.linalg.inv(S)) < 1e-10


def test_phase_split_of_the_default_problem(orc):
    P = orc.Problem(_table(orc, "trot"), 0, 0.6)
    # 1111x11, 1001x20, 0000x5, 0110x19, 0000x5
    assert [(p["contact"], p["horizon"]) for p in P.phases] == [
        ([1, 1, 1, 1], 11), ([1, 0, 0, 1], 20), ([0, 0, 0, 0], 5), ([0, 1, 1, 0], 19), ([0, 0, 0, 0], 5)]
    assert [p["n_td"] for p in P.phases] == [0, 0, 2, 0, 2]
    assert P.n_stages == 60


def test_one_iteration_invariants(orc):
    P = orc.Problem(_table(orc, "trot"), 0, 0.6)
    assert P.hybrid_rollout(0.0)
    P.update_nominal()
    P.compute_cost()
    P.lq_approximation()
    assert P.backward_sweep(0.0)
    s = P.scalars()
    assert s["dV_1"] <= 0 and abs(s["dV_1"] + s["dV_2"]) < 1e-9 * abs(s["dV_1"])
    H = P.get("H")
    assert np.abs(H - np.swapaxes(H, 1, 2)).max() < 1e-10 * np.abs(H).max()
    # K, dU solve the stage-wise stationarity condition: the gains keep Quu positive definite
    K = P.get("K")
    assert np.isfinite(K).all() and np.isfinite(P.get("dU")).all()
    # finite-difference check of the reset-map chain through the defect: with eps = 0 a rollout reproduces Xbar
    X0 = P.get("X").copy()
    assert P.hybrid_rollout(0.0)
    assert np.array_equal(P.get("X"), X0)


def test_sweep_and_linear_rollout_agree_without_defects(orc):
    """With a dynamically consistent nominal trajectory (zero defects) the expected cost change of the
    backward sweep, -sum Qu^T Quu^-1 Qu, equals the one accumulated by linear_rollout(1)."""
    P = orc.Problem(_table(orc, "trot"), 0, 0.25)
    # build a dynamically consistent trajectory: single-shooting rollouts with U = 0, moving each
    # phase's first node (always a shooting node, Q9) onto its propagated initial state
    o = dict(MS=0)
    for _ in range(P.n_phases + 1):
        assert P.hybrid_rollout(0.0, o)
        P.set("Xbar", P.get("X") + P.get("Defect"))
    assert P.hybrid_rollout(0.0, o)
    P.update_nominal()
    assert np.abs(P.get("Defect")).max() == 0.0
    P.compute_cost(o)
    P.lq_approximation(o)
    assert P.backward_sweep(0.0)
    bs = P.scalars()
    P.linear_rollout(1.0, o)
    lr = P.scalars()
    # dV = dV_1 + dV_2/2 is the same quadratic-model decrease in both
    a = bs["dV_1"] + 0.5 * bs["dV_2"]
    b = lr["dV_1"] + 0.5 * lr["dV_2"]
    assert abs(a - b) < 1e-8 * abs(a)


def test_cost_decreases_on_merit(orc):
    P = orc.Problem(_table(orc, "trot"), 0, 0.6)
    s, tr = P.solve()
    assert s["status"] in (0, 1, 2) and s["cost"] < s["cost0"]
    acc = tr[tr[:, 9] > 0]
    merit_before = acc[:, 2] + acc[:, 8] * acc[:, 3]
    merit_after = acc[:, 11] + acc[:, 8] * acc[:, 12]
    assert np.all(merit_after <= merit_before)


def test_matches_committed_golden_records(orc):
    g = np.load(os.path.join(GOLDEN, "oracle_solves.npz"))
    names = sorted({k.split("/")[0] for k in g.files})
    gait_of = lambda n: "bound" if n.startswith("bound") else n.split("_")[0]
    for name in names:
        k0, plan = g[f"{name}/meta"]
        for model in ([orc.MODEL_REF] if orc.ref_available() else []) + [orc.MODEL_PORT]:
            P = orc.Problem(_table(orc, gait_of(name)), int(k0), float(plan), model=model)
            P.x0 = g[f"{name}/x0"]
            s, tr = P.solve()
            gs = g[f"{name}/summary"]
            assert s["n_iter"] == gs[1] and s["status"] == gs[0], name
            assert np.array_equal(tr[:, 9], g[f"{name}/trace"][:, 9]), name          # accepted step sizes
            tol = 0 if model == orc.MODEL_REF else 1e-9
            assert rel_err(tr[:, 11], g[f"{name}/trace"][:, 11]) <= tol, name         # per-iteration cost
            assert rel_err(P.get("Xbar"), g[f"{name}/Xbar"]) <= tol and rel_err(P.get("Ubar"), g[f"{name}/Ubar"]) <= tol, name


def test_survey_probe_numbers(orc):
    """an independent throw-away restatement found the same iteration counts and costs."""
    expect = {("trot", 0, 0.6): (13, 9.3588), ("trot", 0, 0.5): (15, 7.7577), ("trot", 0, 0.25): (18, 4.1499),
              ("trot", 0, 1.0): (11, 21.5325), ("bound", 0, 0.6): (29, 96.2766)}
    for (gait, k0, plan), (iters, cost) in expect.items():
        s, _ = orc.Problem(_table(orc, gait), k0, plan).solve()
        assert s["n_iter"] == iters and abs(s["cost"] - cost) < 1e-4


def test_trial_control_identity_of_the_linearised_rollout(orc):
    """Inside solve() the CUDA trial rollout uses U = (Ubar + eps dU) + eps (K dX) with K dX taken from the linear
    rollout of the same iteration (hybrid_rollout_block<true>), the reference evaluates
    (Ubar + eps dU) + K ((Xbar + eps dX) - Xbar) (SinglePhase.cpp:182-233).  On the oracle's own iterates the two
    differ by rounding only, for every step size of the line search."""
    for gait, plan in (("trot", 0.6), ("bound", 0.5), ("pronk", 0.4)):
        P = orc.Problem(_table(orc, gait), 0, plan)
        assert P.hybrid_rollout(0.0)
        P.update_nominal()
        for it in range(3):
            P.compute_cost()
            P.lq_approximation()
            assert P.backward_sweep(0.0 if it == 0 else 1e-3)
            P.linear_rollout(1.0)
            K, dX, dU, Xbar, Ubar = P.get("K"), P.get("dX"), P.get("dU"), P.get("Xbar"), P.get("Ubar")
            # node of stage s = s + (index of its phase): every phase owns horizon + 1 nodes
            node = np.concatenate([np.arange(ph["horizon"]) + off + i for i, (ph, off) in
                                   enumerate(zip(P.phases, np.cumsum([0] + [p["horizon"] for p in P.phases[:-1]])))])
            KdX = np.einsum("sij,sj->si", K, dX[node])
            for eps in (1.0, 0.1, 0.010000000000000002, 0.0010000000000000002):
                assert P.hybrid_rollout(eps)
                U_ref = P.get("U")
                U_lin = (Ubar + eps * dU) + eps * KdX
                scale = np.abs(K).sum(axis=2) * np.abs(Xbar[node]).max(axis=1, keepdims=True) + np.abs(U_ref) + 1.0
                assert (np.abs(U_lin - U_ref) <= 1e-14 * scale).all(), (gait, it, eps, np.abs(U_lin - U_ref).max())
            assert P.hybrid_rollout(1.0)
            P.update_nominal()


def _solve_ld(M, R):
    """M^-1 R in extended precision (np.longdouble), Gauss-Jordan with partial pivoting: an inverse algorithm that shares
    nothing with oracle/linalg.hpp (LU + LDLT in double)."""
    M = np.array(M, np.longdouble); R = np.array(R, np.longdouble)
    n = M.shape[0]
    T = np.concatenate([M, R.reshape(n, -1)], axis=1)
    for c in range(n):
        p = c + int(np.argmax(np.abs(T[c:, c])))
        T[[c, p]] = T[[p, c]]
        T[c] = T[c] / T[c, c]
        for r in range(n):
            if r != c:
                T[r] = T[r] - T[r, c] * T[c]
    return T[:, n:].reshape(R.shape)


def test_riccati_stage_against_independent_extended_precision_restatement(orc):
    """a second, independent restatement of the Riccati stage (SinglePhase.cpp:307-363) in NumPy
    extended precision, textbook form (dense products, a different inverse algorithm), checked against the oracle's
    K, dU, G, H at every stage of every phase, for a trot and a bound problem after two DDP iterations."""
    for gait, k0 in (("trot", 0), ("bound", 5)):
        P = orc.Problem(_table(orc, gait), k0, 0.6)
        P.solve(dict(max_AL_iter=1, max_DDP_iter=2))
        P.compute_cost(); P.lq_approximation()
        assert P.backward_sweep(0.0)
        ld = np.longdouble
        A, B = P.get("A").astype(ld), P.get("B").astype(ld)
        lx, lu, lxx, luu, lux = (P.get(n).astype(ld) for n in ("lx", "lu", "lxx", "luu", "lux"))
        D, G, H, K, dU = P.get("Defect").astype(ld), P.get("G"), P.get("H"), P.get("K"), P.get("dU")
        s0 = n0 = 0
        worst = 0.0
        for ph in P.phases:
            N = ph["horizon"]
            for k in range(N - 1, -1, -1):
                s, n = s0 + k, n0 + k
                Hn, Gn = H[n + 1].astype(ld), G[n + 1].astype(ld) + H[n + 1].astype(ld) @ D[n + 1]
                Qx, Qu = lx[s] + A[s].T @ Gn, lu[s] + B[s].T @ Gn
                Qxx, Quu, Qux = lxx[s] + A[s].T @ Hn @ A[s], luu[s] + B[s].T @ Hn @ B[s], lux[s] + B[s].T @ Hn @ A[s]
                assert np.all(np.linalg.eigvalsh(np.array(0.5 * (Quu + Quu.T), np.float64)) > 1e-9)
                Kk, dUk = -_solve_ld(Quu, Qux), -_solve_ld(Quu, Qu)
                Gk = Qx + Qux.T @ dUk
                Hk = 0.5 * (Qxx + Qxx.T) + Qux.T @ Kk
                if k == 0:
                    Gk = Gk + Hk @ D[n]  # G[0] += H[0] Defect[0] at the end of the phase (SinglePhase.cpp:365)
                for mine, theirs in ((Kk, K[s]), (dUk, dU[s]), (Gk, G[n]), (Hk, H[n])):
                    e = float(np.abs(np.array(mine, np.float64) - theirs).max() / max(1e-300, np.abs(theirs).max()))
                    worst = max(worst, e)
                    assert e < 1e-10, (gait, s, e)
            s0 += N; n0 += N + 1
        assert worst > 0.0  # (the two paths are different computations: they do not agree bit for bit)
