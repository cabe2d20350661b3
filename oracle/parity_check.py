"""Parity of the CUDA path against the oracle on many problems at once.

TEST INFRASTRUCTURE ONLY (used by tests/ and by the `parity` block of bench.py, where the oracle is the
checker, never the thing measured).  Nothing under hkd-mpc_b200/ imports this.

A DDP solve is a chain of discontinuous decisions (Armijo accept / reject, PD test, termination tests): for some
inputs a perturbation of one unit in the last place flips a decision and the rest of the solve differs at O(1)
(SURVEY.md §7.2).  Such a problem has no "reference answer within 1e-9" — not even between two builds of the
reference itself — so every problem is first CLASSIFIED with the oracle alone, by solving it under up to three
arithmetic variants of the same restatement:
    ref   reference's CasADi model compiled unmodified (oracle/_ref), oracle solver compiled -O3 without contraction
    port  the oracle's independent dual-number model port (differs from `ref` by ~1e-16 per model call)
    fma   the `ref` variant compiled with FMA contraction (-mfma -ffp-contract=fast: what a -march=native build of
          the reference would do to its own arithmetic)
A problem is WELL-POSED for parity when all variants take the same decisions (status, iterations, outer iterations,
backward sweeps, line-search trials) and agree to `sens_tol` on the final cost and trajectories.  On well-posed
problems the CUDA path must take exactly the reference variant's decisions and agree to `rtol` = 1e-9
(BASELINE.json north_star); ill-posed ones are counted and reported, not compared.
"""
import numpy as np

RTOL = 1e-9
SENS_TOL = 1e-11
KEYS = ("status", "n_iter", "n_outer", "n_sweeps")


def rel_err_rows(a, b, floor=1e-6):
    """Per-problem relative error with a per-ROW denominator: max over rows r of max|a_r - b_r| / max(max|b_r|, floor * max|b|).
    a, b: [n, rows, ...].  A plain max-norm over the whole array would hide errors in small rows (e.g. gain rows of
    weakly coupled controls); the floor keeps structurally-zero rows from dividing by zero."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    n, rows = a.shape[0], a.shape[1]
    d = np.abs(a - b).reshape(n, rows, -1).max(axis=2)
    m = np.abs(b).reshape(n, rows, -1).max(axis=2)
    g = np.maximum(m.max(axis=1, keepdims=True), 1e-300)
    return (d / np.maximum(m, floor * g)).max(axis=1)


def _decisions(r):
    s = r["summary"]
    return np.stack([s[:, 0], s[:, 1], s[:, 2], s[:, 3], r["n_trials"].astype(np.float64)], axis=1)


def _agreement(r, q):
    """(same decisions [n] bool, sensitivity [n]) between two oracle runs."""
    same = np.all(_decisions(r) == _decisions(q), axis=1)
    c0, c1 = r["summary"][:, 4], q["summary"][:, 4]
    sens = np.abs(c0 - c1) / np.maximum(np.abs(c0), 1e-300)
    sens = np.maximum(sens, rel_err_rows(q["Xbar"], r["Xbar"]))
    sens = np.maximum(sens, rel_err_rows(q["Ubar"], r["Ubar"]))
    if r.get("K") is not None:
        sens = np.maximum(sens, rel_err_rows(q["K"], r["K"]))
    return same, np.where(same, sens, np.inf)


def oracle_runs(orc, tables, k0, x0, max_nodes, max_stages, plan, k_rows=8, opts=None, cparams=None, n_threads=0, variants=("port", "fma")):
    """Solve the problems with the reference variant (timed: this is also bench.py's CPU sample) and the requested
    perturbed variants.  Returns (ref_run, {name: run})."""
    base = orc.batch_solve_traj(tables, k0, x0, max_nodes, max_stages, k_rows, plan=plan, opts=opts, cparams=cparams, n_threads=n_threads)
    others = {}
    for v in variants:
        if v == "port":
            if not orc.ref_available():
                continue  # the base run already is the port
            others[v] = orc.batch_solve_traj(tables, k0, x0, max_nodes, max_stages, k_rows, plan=plan, model=orc.MODEL_PORT, opts=opts,
                                             cparams=cparams, n_threads=n_threads)
        elif v == "fma":
            if not orc.variant_available("fma"):
                continue
            others[v] = orc.batch_solve_traj(tables, k0, x0, max_nodes, max_stages, k_rows, plan=plan, opts=opts, cparams=cparams,
                                             n_threads=n_threads, variant="fma")
    return base, others


def classify(base, others, sens_tol=SENS_TOL):
    """well_posed [n] bool and the per-problem sensitivity (max over variants)."""
    n = base["summary"].shape[0]
    well = np.ones(n, bool)
    sens = np.zeros(n)
    for q in others.values():
        same, s = _agreement(base, q)
        well &= same & (s < sens_tol)
        sens = np.maximum(sens, s)
    return well, sens


def compare_gpu(base, well, gpu, rtol=RTOL):
    """gpu: dict(info [n] structured hsddp_info, Xbar [n,max_nodes,24], Ubar [n,max_stages,24], K [n,k_rows,24,24] or None).
    Returns the report dict (JSON-serialisable) and the per-problem arrays (match, err)."""
    info = gpu["info"]
    s = base["summary"]
    dec_ok = ((info["status"] == s[:, 0]) & (info["n_iter"] == s[:, 1]) & (info["n_outer"] == s[:, 2]) & (info["n_sweeps"] == s[:, 3])
              & (info["n_trials"] == base["n_trials"]))
    e_cost = np.abs(info["cost"] - s[:, 4]) / np.maximum(np.abs(s[:, 4]), 1e-300)
    e_x = rel_err_rows(gpu["Xbar"], base["Xbar"])
    e_u = rel_err_rows(gpu["Ubar"], base["Ubar"])
    e_k = rel_err_rows(gpu["K"], base["K"]) if gpu.get("K") is not None and base.get("K") is not None else np.zeros(len(s))
    err = np.maximum(np.maximum(e_cost, e_x), np.maximum(e_u, e_k))
    err = np.where(np.isfinite(err), err, np.inf)
    match = dec_ok & (err < rtol)
    w = well
    nw = int(w.sum())
    rep = {
        "checked": int(len(s)), "well_posed": nw, "ill_posed": int(len(s) - nw),
        "iter_status_match": int((dec_ok & w).sum()),          # same status / iterations / outer / sweeps / trials as the oracle
        "within_1e-9": int((match & w).sum()),                  # ... and cost, Xbar, Ubar, K (per-row relative) within rtol
        "max_rel_cost": float(e_cost[w & dec_ok].max()) if (w & dec_ok).any() else None,
        "max_rel_Xbar": float(e_x[w & dec_ok].max()) if (w & dec_ok).any() else None,
        "max_rel_Ubar": float(e_u[w & dec_ok].max()) if (w & dec_ok).any() else None,
        "max_rel_K": float(e_k[w & dec_ok].max()) if (w & dec_ok).any() else None,
        "ill_posed_also_matching": int((match & ~w).sum()),
        "rtol": rtol,
    }
    return rep, dict(match=match, dec_ok=dec_ok, err=err)
