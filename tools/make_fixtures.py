#!/usr/bin/env python3
"""Generate the committed gait tables under hkd-mpc_b200/data/ from the reference tree.

Runs ONLY in the build container (needs /root/reference).  Nothing at test or
bench time reads /root/reference: the GPU box gets these fixtures instead.

gait_<name>.npz — what `QuadReference::load_top_level_data`
(Reference/QuadReference.cpp:129-285) leaves in `tp_data` after reading a
`quad_reference.csv`: every number passed through std::stof, so float32 is
exact.  Keys are matched by substring in the loader's fixed
order; files that use `jnt_angle` leave qJ at zero (Q12).

  trot   <- Reference/Data/trot/quad_reference.csv
  bound  <- scripts/ReferenceGen/PostProcessedData/quad_reference.csv
  pronk  <- scripts/ReferenceGen/PreProcessedData/MixedHopping/*.csv converted by
            the rule of scripts/ReferenceGen/generate_reference.m:7-57 (GRF =
            9 kg * 10 / n_contacts on z of each stance leg, values printed with
            3 decimals then parsed like the others)
"""
import os
import sys
import numpy as np

REF = os.environ.get("HKD_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hkd-mpc_b200", "data")


def load_quad_reference(path):
    """Restates QuadReference::load_top_level_data line by line."""
    body, qJ, foot, grf, contact, sdur = [], [], [], [], [], []
    dt = None
    cur = None
    with open(path) as f:
        lines = f.read().split("\n")
    i = 0

    def floats(line, n):
        w = line.split()
        v = np.zeros(n, np.float32)
        for j, s in enumerate(w[:n]):
            v[j] = np.float32(s)
        return v

    while i < len(lines):
        line = lines[i]
        i += 1
        if line == "dt":
            dt = np.float32(lines[i]); i += 1
            continue
        if "body_state" in line:
            cur = dict(body=floats(lines[i], 12), qJ=np.zeros(12, np.float32), foot=np.zeros(12, np.float32),
                       grf=np.zeros(12, np.float32), contact=np.zeros(4, np.int32), sdur=np.zeros(4, np.float32))
            i += 1
            continue
        if "qJ" in line:
            cur["qJ"] = floats(lines[i], 12); i += 1
            continue
        if "foot_placements" in line:
            cur["foot"] = floats(lines[i], 12); i += 1
            continue
        if "grf" in line:
            cur["grf"] = floats(lines[i], 12); i += 1
            continue
        if "torque" in line:
            i += 1
            continue
        if "contact" in line:
            w = lines[i].split(); i += 1
            cur["contact"] = np.array([int(s) for s in w[:4]], np.int32)
            continue
        if "status_dur" in line:
            cur["sdur"] = floats(lines[i], 4); i += 1
            body.append(cur["body"]); qJ.append(cur["qJ"]); foot.append(cur["foot"]); grf.append(cur["grf"])
            contact.append(cur["contact"]); sdur.append(cur["sdur"])
    return dict(dt=np.float32(dt), body_state=np.stack(body), qJ=np.stack(qJ), foot_placements=np.stack(foot),
                grf=np.stack(grf), contact=np.stack(contact), status_dur=np.stack(sdur))


def fmt3(a):
    """'%6.3f' print then std::stof, as generate_reference.m + the loader do."""
    return np.array([[np.float32("%6.3f" % v) for v in row] for row in a], np.float32)


def convert_preprocessed(folder):
    """generate_reference.m:7-57 for a PreProcessedData gait folder."""
    body = np.loadtxt(os.path.join(folder, "body_state.csv"), delimiter=",")
    contacts = np.loadtxt(os.path.join(folder, "contact.csv"), delimiter=",").astype(np.int32)
    foot = np.loadtxt(os.path.join(folder, "ee_pos.csv"), delimiter=",")
    qJ = np.loadtxt(os.path.join(folder, "jnt.csv"), delimiter=",")
    t = np.loadtxt(os.path.join(folder, "time.csv"), delimiter=",")
    n = body.shape[0]
    dt = t[1] - t[0]
    grf = np.zeros((n, 12))
    mass, g = 9.0, 10.0
    for k in range(n):
        nc = contacts[k].sum()
        for leg in range(4):
            if contacts[k, leg]:
                grf[k, 3 * leg + 2] = mass * g / nc
    # status durations (Induce_status_duration_per_leg)
    sdur = np.zeros((n, 4))
    for leg in range(4):
        c = contacts[:, leg]
        status_dur, start, prev = 0.0, 0, c[0]
        for k in range(1, n):
            status_dur += dt
            if c[k] != prev:
                sdur[start:k, leg] = status_dur
                start, status_dur, prev = k, 0.0, c[k]
            if k == n - 1:
                sdur[start:k + 1, leg] = status_dur
    return dict(dt=np.float32("%4.3f" % dt), body_state=fmt3(body), qJ=fmt3(qJ), foot_placements=fmt3(foot),
                grf=fmt3(grf), contact=contacts, status_dur=fmt3(sdur))


def main():
    os.makedirs(OUT, exist_ok=True)
    gaits = {
        "trot": load_quad_reference(os.path.join(REF, "Reference/Data/trot/quad_reference.csv")),
        "bound": load_quad_reference(os.path.join(REF, "scripts/ReferenceGen/PostProcessedData/quad_reference.csv")),
        "pronk": convert_preprocessed(os.path.join(REF, "scripts/ReferenceGen/PreProcessedData/MixedHopping")),
    }
    for name, g in gaits.items():
        path = os.path.join(OUT, f"gait_{name}.npz")
        np.savez_compressed(path, **g)
        print(name, g["body_state"].shape, "dt", g["dt"], os.path.getsize(path), "bytes")
    # self-check of the conversion rule: the reference's own PostProcessed file was produced by
    # generate_reference.m from PreProcessedData/RunJump*, so converting that folder must reproduce it.
    for cand in ("RunJump", "RunJump_ICRA23"):
        conv = convert_preprocessed(os.path.join(REF, "scripts/ReferenceGen/PreProcessedData", cand))
        b = gaits["bound"]
        same = all(conv[k].shape == b[k].shape and np.array_equal(conv[k], b[k])
                   for k in ("body_state", "qJ", "foot_placements", "grf", "contact"))
        print(f"conversion rule reproduces PostProcessedData from {cand}: {same}")


if __name__ == "__main__":
    sys.exit(main())
