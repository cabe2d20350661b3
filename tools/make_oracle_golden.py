#!/usr/bin/env python3
"""Golden solve records produced by the CPU oracle running on the REFERENCE's own
compiled CasADi model (oracle/_ref).  They are a regression pin for the oracle and a
committed fixture for the GPU parity tests (tests/golden/oracle_solves.npz).  The
reference ships no recorded outputs for this path, so these are
oracle-generated, not reference-generated: "parity unpinned" at the solver level."""
import importlib
import os
import sys
import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as op  # noqa: E402

CASES = [  # (name, gait, window_start, plan, x0 mode)
    ("trot_0.6", "trot", 0, 0.6, "default"),
    ("trot_0.5", "trot", 0, 0.5, "default"),
    ("trot_0.25", "trot", 0, 0.25, "default"),
    ("trot_1.0", "trot", 0, 1.0, "default"),
    ("bound_0.6", "bound", 0, 0.6, "default"),
    ("boundjump_250", "bound", 250, 0.6, "reference"),
    ("pronk_100", "pronk", 100, 0.6, "reference"),
]


def main():
    assert op.ref_available()
    out = {}
    for name, gait, k0, plan, mode in CASES:
        T = op.GaitTable(os.path.join(ROOT, "hkd-mpc_b200", "data", f"gait_{gait}.npz"))
        P = op.Problem(T, k0, plan, model=op.MODEL_REF)
        if mode == "reference":
            body = T.body_state[k0].astype(np.float64)
            qd = op.model_hkd_state(op.MODEL_REF, body[0:3], body[3:6], np.array([0, -0.8, 1.6] * 4, np.float64), np.array(P.phases[0]["contact"], np.int32))
            P.x0 = np.concatenate([body, qd])
        x0 = P.x0
        s, tr = P.solve()
        out[f"{name}/meta"] = np.array([k0, plan])
        out[f"{name}/x0"] = x0
        out[f"{name}/summary"] = np.array([s[k] for k in op.SUMMARY_COLS])
        out[f"{name}/trace"] = tr
        out[f"{name}/Xbar"] = P.get("Xbar")
        out[f"{name}/Ubar"] = P.get("Ubar")
        out[f"{name}/dU"] = P.get("dU")
        out[f"{name}/K_first8"] = P.get("K")[:8]
        print(name, {k: s[k] for k in ("status", "n_iter", "cost")}, "eps", tr[:, 9])
    path = os.path.join(ROOT, "tests", "golden", "oracle_solves.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
