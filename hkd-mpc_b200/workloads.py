"""Synthetic benchmark / test workloads of SURVEY.md §8(d) (BASELINE.json `configs`).

Every workload is a list of schedules (reference windows) + a schedule id and an
initial state per problem.  Initial-state perturbations follow §8(d) config 2:
uniform in eul +-0.05 rad, pos +-0.02 m, omega +-0.2 rad/s, vel +-0.1 m/s on the
body states only, splitmix64 seeded 0xB200 + i, 12 draws per problem; qdummy is
recomputed by the compute_hkd_state rule (HKDModel.h:65-96) with qJ unchanged.
The gait tables come from tests/golden/gait_*.npz (tools/make_fixtures.py).
"""
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(_HERE, "..", "tests", "golden")
GAITS = ("trot", "bound", "pronk")
DEFAULT_QJ = np.array([0, -0.8, 1.6] * 4, np.float64)  # HKDMPC.cpp:47
_MASK = (1 << 64) - 1
_AMP = np.array([0.05] * 3 + [0.02] * 3 + [0.2] * 3 + [0.1] * 3)


def splitmix64_uniform(seed, n):
    """n doubles in [0,1) from splitmix64 (53 high bits)."""
    out = np.empty(n)
    s = seed & _MASK
    for i in range(n):
        s = (s + 0x9E3779B97F4A7C15) & _MASK
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        z = z ^ (z >> 31)
        out[i] = (z >> 11) * (1.0 / 9007199254740992.0)
    return out


def perturbation(i):
    """delta_i on the 12 body states for problem index i."""
    return (2.0 * splitmix64_uniform(0xB200 + i, 12) - 1.0) * _AMP


def gait_path(name):
    return os.path.join(GOLDEN, f"gait_{name}.npz")


class Workload:
    def __init__(self, name, schedules, schedule_id, x0, keys, plan):
        self.name = name
        self.schedules = schedules          # list of pkg.Schedule
        self.schedule_id = np.asarray(schedule_id, np.int32)
        self.x0 = np.ascontiguousarray(x0, np.float64)
        self.keys = keys                    # per schedule: (gait, window_start)
        self.plan = plan
        self.n = len(self.schedule_id)


def _build(pkg, name, entries, plan, perturb_from):
    """entries: per problem (gait, k0, x0_mode) ; x0_mode 'default' or 'reference'."""
    refs = {}
    tables = {}
    sched_index = {}
    schedules, keys, sid = [], [], []
    x0 = np.zeros((len(entries), 24))
    for i, (gait, k0, mode) in enumerate(entries):
        if gait not in refs:
            refs[gait] = pkg.QuadReference(gait_path(gait))
            tables[gait] = np.load(gait_path(gait))
        key = (gait, int(k0))
        if key not in sched_index:
            sched_index[key] = len(schedules)
            schedules.append(pkg.Schedule(refs[gait], int(k0), plan))
            keys.append(key)
        s = schedules[sched_index[key]]
        sid.append(sched_index[key])
        if mode == "default":
            body = s.default_x0()[:12].copy()
        else:
            body = tables[gait]["body_state"][int(k0)].astype(np.float64)
        if perturb_from is not None and i >= perturb_from:
            body = body + perturbation(i)
        x0[i, :12] = body
        x0[i, 12:] = pkg.compute_hkd_state(body[0:3], body[3:6], DEFAULT_QJ, s.contact[0])
    return Workload(name, schedules, sid, x0, keys, plan)


def config1(pkg, plan=0.6):
    """single HKD-MPC solve, Mini Cheetah trot, window 0, x0 of HKDMPC.cpp:44-54."""
    return _build(pkg, "config1: single trot solve", [("trot", 0, "default")], plan, None)


def config2(pkg, n=1024, plan=0.6):
    """n trot problems with perturbed initial body states (problem 0 unperturbed)."""
    return _build(pkg, f"config2: {n} trot problems, perturbed x0", [("trot", 0, "default")] * n, plan, 1)


def config3(pkg, n=16384, plan=0.6, first=0):
    """n mixed-gait problems: gait i mod 3, window start (7*(i div 3)) mod (n_samples-63),
    x0 = reference body state at the window start + perturbation.  `first` offsets the
    problem index (used to shard by index across ranks)."""
    sizes = {g: np.load(gait_path(g))["body_state"].shape[0] for g in GAITS}
    entries = []
    for j in range(n):
        i = first + j
        g = GAITS[i % 3]
        k0 = (7 * (i // 3)) % (sizes[g] - (int(round(plan / 0.01)) + 3))  # (n_samples - 63 for the 0.6 s horizon of SURVEY.md §8d)
        entries.append((g, k0, "reference"))
    w = _build_indexed(pkg, f"config3: {n} mixed-gait problems (trot/bound/pronk)", entries, plan, first)
    return w


def config4(pkg, n=4096, plan=0.6):
    """n bound-then-jump problems, window starts uniform in [236, 266] (long flight phase in the horizon)."""
    entries = []
    for i in range(n):
        k0 = 236 + int(splitmix64_uniform(0xB200 + (1 << 20) + i, 1)[0] * 31)
        entries.append(("bound", min(k0, 266), "reference"))
    return _build(pkg, f"config4: {n} bound+jump problems", entries, plan, 0)


def _build_indexed(pkg, name, entries, plan, first):
    """like _build with perturbation index = global problem index."""
    refs, tables, sched_index = {}, {}, {}
    schedules, keys, sid = [], [], []
    x0 = np.zeros((len(entries), 24))
    for j, (gait, k0, mode) in enumerate(entries):
        if gait not in refs:
            refs[gait] = pkg.QuadReference(gait_path(gait))
            tables[gait] = np.load(gait_path(gait))
        key = (gait, int(k0))
        if key not in sched_index:
            sched_index[key] = len(schedules)
            schedules.append(pkg.Schedule(refs[gait], int(k0), plan))
            keys.append(key)
        s = schedules[sched_index[key]]
        sid.append(sched_index[key])
        body = tables[gait]["body_state"][int(k0)].astype(np.float64) + perturbation(first + j)
        x0[j, :12] = body
        x0[j, 12:] = pkg.compute_hkd_state(body[0:3], body[3:6], DEFAULT_QJ, s.contact[0])
    return Workload(name, schedules, sid, x0, keys, plan)
