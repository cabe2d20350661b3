"""Synthetic benchmark / test workloads of SURVEY.md §8(d) (BASELINE.json `configs`).

Every workload is a list of schedules (reference windows) + a schedule id and an
initial state per problem.  Initial-state perturbations follow §8(d) config 2:
uniform in eul +-0.05 rad, pos +-0.02 m, omega +-0.2 rad/s, vel +-0.1 m/s on the
body states only, splitmix64 seeded 0xB200 + i, 12 draws per problem; qdummy is
recomputed by the compute_hkd_state rule (HKDModel.h:65-96) with qJ unchanged.
The gait tables are hkd-mpc_b200/data/gait_*.npz (generated from the reference's
data files by tools/make_fixtures.py).

The workload DEFINITION (which gait, which window, which initial state) is pure
NumPy: `entries_*` + `initial_states` need only a compute_hkd_state function, so
the CPU arm of bench.py builds the identical inputs without loading the CUDA
library.  `config1..4(pkg, ...)` add the flattened schedules for the GPU path.
"""
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(_HERE, "data")
GAITS = ("trot", "bound", "pronk")
DEFAULT_QJ = np.array([0, -0.8, 1.6] * 4, np.float64)  # HKDMPC.cpp:47
DEFAULT_BODY = np.array([0, 0, 0, 0, 0, 0.2486, 0, 0, 0, 0, 0, 0], np.float64)  # HKDMPC.cpp:44-46
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
    return os.path.join(DATA, f"gait_{name}.npz")


_tables = {}


def gait_table(name):
    if name not in _tables:
        d = np.load(gait_path(name))
        _tables[name] = {k: d[k] for k in d.files}
    return _tables[name]


# ---------------------------------------------------------------------------
# workload definitions: one entry per problem = (gait, window start, x0 mode, perturbation index or None)
# ---------------------------------------------------------------------------
def entries_config1():
    return [("trot", 0, "default", None)]


def entries_config2(n):
    return [("trot", 0, "default", (i if i >= 1 else None)) for i in range(n)]


def entries_config3(n, plan=0.6, first=0, stride=1):
    sizes = {g: gait_table(g)["body_state"].shape[0] for g in GAITS}
    out = []
    for j in range(n):
        i = first + j * stride
        g = GAITS[i % 3]
        k0 = (7 * (i // 3)) % (sizes[g] - (int(round(plan / 0.01)) + 3))  # (n_samples - 63 for the 0.6 s horizon of SURVEY.md §8d)
        out.append((g, int(k0), "reference", i))
    return out


def entries_config4(n):
    out = []
    for i in range(n):
        k0 = 236 + int(splitmix64_uniform(0xB200 + (1 << 20) + i, 1)[0] * 31)
        out.append(("bound", min(k0, 266), "reference", i))
    return out


def initial_states(entries, hkd_state):
    """x0 [n, 24] of the entries.  hkd_state(eul, pos, qJ, contact[4]) -> qdummy[12] (compute_hkd_state, HKDModel.h:65-96);
    the contact is that of the first phase = the reference sample at the window start."""
    x0 = np.zeros((len(entries), 24))
    for j, (gait, k0, mode, pidx) in enumerate(entries):
        t = gait_table(gait)
        body = DEFAULT_BODY.copy() if mode == "default" else t["body_state"][k0].astype(np.float64)
        if pidx is not None:
            body = body + perturbation(pidx)
        x0[j, :12] = body
        x0[j, 12:] = hkd_state(body[0:3], body[3:6], DEFAULT_QJ, np.asarray(t["contact"][k0], np.int32))
    return x0


def schedule_keys(entries):
    """Deduplicated (gait, window start) pairs and the schedule id of every entry."""
    index, keys, sid = {}, [], []
    for gait, k0, _, _ in entries:
        key = (gait, int(k0))
        if key not in index:
            index[key] = len(keys)
            keys.append(key)
        sid.append(index[key])
    return keys, np.asarray(sid, np.int32)


class Workload:
    def __init__(self, name, schedules, schedule_id, x0, keys, plan, entries=None):
        self.name = name
        self.schedules = schedules          # list of pkg.Schedule (None for the CPU-only form)
        self.schedule_id = np.asarray(schedule_id, np.int32)
        self.x0 = np.ascontiguousarray(x0, np.float64)
        self.keys = keys                    # per schedule: (gait, window_start)
        self.plan = plan
        self.entries = entries
        self.n = len(self.schedule_id)


NAMES = {"config1": "config1: single trot solve", "config2": "config2: {n} trot problems, perturbed x0",
         "config3": "config3: {n} mixed-gait problems (trot/bound/pronk)", "config4": "config4: {n} bound+jump problems"}


def define(config, n=None, plan=0.6, first=0, stride=1):
    """(name, entries) of a configuration — no library needed."""
    if config == "config1":
        e = entries_config1()
    elif config == "config2":
        e = entries_config2(n)
    elif config == "config3":
        e = entries_config3(n, plan, first, stride)
    elif config == "config4":
        e = entries_config4(n)
    else:
        raise ValueError(config)
    return NAMES[config].format(n=len(e)), e


def build_cpu(config, hkd_state, n=None, plan=0.6, first=0, stride=1):
    """The workload without schedules (CPU arm / oracle side): keys, schedule ids, x0."""
    name, e = define(config, n, plan, first, stride)
    keys, sid = schedule_keys(e)
    return Workload(name, None, sid, initial_states(e, hkd_state), keys, plan, e)


def _build(pkg, config, n=None, plan=0.6, first=0, stride=1):
    name, e = define(config, n, plan, first, stride)
    return from_entries(pkg, name, e, plan)


def from_entries(pkg, name, e, plan=0.6):
    """A workload from explicit entries (gait, window start, x0 mode, perturbation index)."""
    keys, sid = schedule_keys(e)
    refs = {}
    schedules = []
    for gait, k0 in keys:
        if gait not in refs:
            refs[gait] = pkg.QuadReference(gait_path(gait))
        schedules.append(pkg.Schedule(refs[gait], int(k0), plan))
    return Workload(name, schedules, sid, initial_states(e, pkg.compute_hkd_state), keys, plan, e)


def config1(pkg, plan=0.6):
    """single HKD-MPC solve, Mini Cheetah trot, window 0, x0 of HKDMPC.cpp:44-54."""
    return _build(pkg, "config1", None, plan)


def config2(pkg, n=1024, plan=0.6):
    """n trot problems with perturbed initial body states (problem 0 unperturbed)."""
    return _build(pkg, "config2", n, plan)


def config3(pkg, n=16384, plan=0.6, first=0, stride=1):
    """n mixed-gait problems: gait i mod 3, window start (7*(i div 3)) mod (n_samples-63),
    x0 = reference body state at the window start + perturbation.  Problem j of the workload is problem
    i = first + j * stride of the configuration (used to shard by index across ranks: contiguous ranges or interleaved)."""
    return _build(pkg, "config3", n, plan, first, stride)


def config4(pkg, n=4096, plan=0.6):
    """n bound-then-jump problems, window starts uniform in [236, 266] (long flight phase in the horizon)."""
    return _build(pkg, "config4", n, plan)


def with_room_for_ticks(entries, ticks, plan=0.6):
    """The entries whose reference table is long enough for `ticks` receding-horizon updates."""
    need = int(round(plan / 0.01)) + 3 + ticks
    return [e for e in entries if e[1] + need < gait_table(e[0])["body_state"].shape[0]]


def gait_batch(pkg, w, device=0, cparams=None):
    """The workload `w` as a batch whose schedules are built ON THE DEVICE from the gait library
    (hsddp_batch_set_problems_from_gaits): the form that supports the receding-horizon update (mpc_update)."""
    order = list(GAITS)
    refs = [pkg.QuadReference(gait_path(g)) for g in order]
    B = pkg.MultiPhaseDDPBatch(device)
    B.set_problems_from_gaits(refs, [order.index(g) for g, _ in w.keys], [k for _, k in w.keys], w.plan, w.schedule_id, cparams)
    B.set_initial_condition(w.x0)
    return B


def random_phase(xs, us, ys, N, seed, n=None):
    """Synthetic plug-in outputs of one phase (or n phases) of SinglePhase<double, xs, us, ys> for the generic sweeps
    (the reference ships no model for <12,12,0> and <36,12,12>): near-identity dynamics, positive definite
    cost Hessians, small defects.  Matrices as [..., row, col]."""
    rng = np.random.default_rng(seed)
    lead = () if n is None else (n,)

    def pd(m, shape):
        M = rng.normal(size=shape + (m, m)) * 0.3
        return M @ np.swapaxes(M, -1, -2) + np.eye(m)
    return dict(A=np.eye(xs) + 0.1 * rng.normal(size=lead + (N, xs, xs)), B=0.3 * rng.normal(size=lead + (N, xs, us)),
                C=0.5 * rng.normal(size=lead + (N, ys, xs)), D=0.5 * rng.normal(size=lead + (N, ys, us)),
                lx=rng.normal(size=lead + (N, xs)), lu=rng.normal(size=lead + (N, us)), ly=rng.normal(size=lead + (N, ys)),
                lxx=pd(xs, lead + (N,)), luu=pd(us, lead + (N,)), lux=0.1 * rng.normal(size=lead + (N, us, xs)),
                lyy=pd(ys, lead + (N,)) if ys else np.zeros(lead + (N, 0, 0)), Phix=rng.normal(size=lead + (xs,)),
                Phixx=pd(xs, lead), Defect=0.05 * rng.normal(size=lead + (N + 1, xs)))


def generic_flop_per_stage(xs, us, ys):
    """Dense algorithmic FLOP of one Riccati stage with product reuse and Cholesky solves (for any sizes)."""
    f = 2 * xs * xs + 2 * xs * (xs + us) + 2 * xs * xs * (xs + us) + 2 * xs ** 3 + 2 * us * xs * xs + 2 * us * us * xs + us ** 3 // 3 \
        + 2 * us * us * (xs + 1) + 2 * xs * us + 2 * xs * xs * us
    if ys:
        f += 2 * xs * ys * ys + 2 * us * ys * ys + 2 * xs * xs * ys + 2 * us * xs * ys + 2 * us * us * ys + 2 * xs * ys + 2 * us * ys
    return f


def generic_bytes_per_stage(xs, us, ys):
    rd = xs * xs + xs * us + ys * xs + ys * us + xs + us + ys + xs * xs + us * us + us * xs + ys * ys + xs
    wr = us + us * xs + xs + xs * xs
    return 8 * (rd + wr)
