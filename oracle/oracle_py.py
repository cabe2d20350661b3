"""ctypes driver for the CPU oracle (oracle/liboracle_hsddp.so).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by hkd-mpc_b200/.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_hsddp.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libhkd_casadi_ref.so")

# HSDDP_OPTION as consumed by solve() — HKDMPC/settings/ddp_setting.info with
# update_regularization = 2 (never loaded).
OPTION_FIELDS = ["alpha", "gamma", "update_penalty", "update_relax", "update_regularization", "update_ReB",
                 "max_DDP_iter", "max_AL_iter", "cost_thresh", "tconstr_thresh", "pconstr_thresh",
                 "dynamics_feas_thresh", "merit_scale", "merit_offset", "AL_active", "ReB_active", "MS"]
DEFAULT_OPTIONS = dict(alpha=0.1, gamma=0.01, update_penalty=5, update_relax=1, update_regularization=2, update_ReB=1,
                       max_DDP_iter=10, max_AL_iter=5, cost_thresh=1e-3, tconstr_thresh=1e-3, pconstr_thresh=1e-3,
                       dynamics_feas_thresh=1e-3, merit_scale=0.2, merit_offset=1e2, AL_active=1, ReB_active=1, MS=1)
# HKDMPC/settings/constraint_params.info + mu (HKDConstraints.h:17)
CPARAM_FIELDS = ["grf_delta", "grf_delta_min", "grf_eps", "td_sigma", "td_sigma_max", "td_lambda", "mu"]
DEFAULT_CPARAMS = dict(grf_delta=0.1, grf_delta_min=0.1, grf_eps=0.1, td_sigma=50.0, td_sigma_max=1e4, td_lambda=0.0, mu=0.7)

MODEL_REF, MODEL_PORT = 0, 1

TRACE_COLS = ["outer", "inner", "cost_before", "feas_before", "reg_after", "n_sweeps", "dV_1", "dV_2", "merit_rho",
              "eps_accepted", "n_trials", "cost_after", "feas_after", "max_tconstr", "max_pconstr", "pad"]
SUMMARY_COLS = ["status", "n_iter", "n_outer", "n_sweeps", "cost", "feas", "max_tconstr", "max_pconstr", "cost0", "feas0"]

_lib = None


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))):
        subprocess.check_call(["make", "-C", _HERE, "liboracle_hsddp.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference") and not os.path.exists(_REF_PATH):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_table_create.restype = C.c_void_p
        L.orc_table_create.argtypes = [C.c_int, C.c_float] + [C.POINTER(C.c_float)] * 4 + [C.POINTER(C.c_int)]
        L.orc_table_destroy.argtypes = [C.c_void_p]
        L.orc_problem_create.restype = C.c_void_p
        L.orc_problem_create.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.POINTER(C.c_double)]
        L.orc_problem_phase_flags.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 4
        L.orc_problem_window_start.argtypes = [C.c_void_p]
        L.orc_problem_window_start.restype = C.c_int
        for f in ("orc_problem_destroy", "orc_update_nominal", "orc_mpc_update"):
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("orc_problem_n_phases", "orc_problem_n_stages"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int
        L.orc_problem_phase_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                             C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_problem_set_x0.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.orc_problem_get_x0.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.orc_problem_stage_reference.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 4 + [C.POINTER(C.c_int)]
        L.orc_problem_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.orc_problem_set.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.orc_problem_scalars.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.orc_hybrid_rollout.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
        L.orc_hybrid_rollout.restype = C.c_int
        L.orc_compute_cost.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.orc_lq_approximation.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.orc_backward_sweep.argtypes = [C.c_void_p, C.c_double]
        L.orc_backward_sweep.restype = C.c_int
        L.orc_linear_rollout.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
        L.orc_solve.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        L.orc_solve.restype = C.c_int
        L.orc_batch_solve.restype = C.c_double
        L.orc_batch_solve.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int, C.c_float,
                                      C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        L.orc_batch_solve_traj.restype = C.c_double
        L.orc_batch_solve_traj.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int, C.c_float,
                                           C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double),
                                           C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int,
                                           C.POINTER(C.c_int)]
        L.orc_load_ref.argtypes = [C.c_char_p]
        L.orc_load_ref.restype = C.c_int
        L.orc_model_dynamics.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_model_dynamics_partial.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_int),
                                                 C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_model_foot_position.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 3 + [C.c_int, C.POINTER(C.c_double)]
        L.orc_model_foot_jacobian.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 3 + [C.c_int, C.POINTER(C.c_double)]
        L.orc_model_hkd_state.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_ldlt_is_positive.argtypes = [C.POINTER(C.c_double)]
        L.orc_ldlt_is_positive.restype = C.c_int
        L.orc_inverse.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
        if os.path.exists(_REF_PATH):
            L.orc_load_ref(_REF_PATH.encode())
        _lib = L
    return _lib


_variants = {}


def _variant_path(name):
    return os.path.join(_HERE, f"liboracle_hsddp_{name}.so")


def variant_available(name):
    """Other builds of the same restatement (oracle/Makefile): "fma" = FMA-contracted arithmetic."""
    try:
        return _variant_lib(name) is not None
    except Exception:
        return False


def _variant_lib(name):
    if name in (None, "", "base"):
        return lib()
    if name not in _variants:
        lib()
        path = _variant_path(name)
        srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
        if not os.path.exists(path) or any(os.path.getmtime(f) > os.path.getmtime(path) for f in srcs):
            subprocess.check_call(["make", "-C", _HERE, os.path.basename(path)], stdout=subprocess.DEVNULL)
        L = C.CDLL(path)
        L.orc_batch_solve_traj.restype = C.c_double
        L.orc_batch_solve_traj.argtypes = lib().orc_batch_solve_traj.argtypes
        L.orc_load_ref.argtypes = [C.c_char_p]
        L.orc_load_ref.restype = C.c_int
        if os.path.exists(_REF_PATH):
            L.orc_load_ref(_REF_PATH.encode())
        _variants[name] = L
    return _variants[name]


def ref_available():
    return bool(lib().orc_ref_loaded())


def default_model():
    """The reference's own compiled CasADi model when present, else the port."""
    return MODEL_REF if ref_available() else MODEL_PORT


def options_array(**over):
    o = dict(DEFAULT_OPTIONS)
    o.update(over)
    return np.array([float(o[k]) for k in OPTION_FIELDS], np.float64)


def cparams_array(**over):
    c = dict(DEFAULT_CPARAMS)
    c.update(over)
    return np.array([float(c[k]) for k in CPARAM_FIELDS], np.float64)


# ---------------- model-level helpers ----------------
def model_dynamics(kind, x, u, dt, c):
    x = np.ascontiguousarray(x, np.float64); u = np.ascontiguousarray(u, np.float64)
    c = np.ascontiguousarray(c, np.int32); xn = np.zeros(24)
    lib().orc_model_dynamics(kind, _dp(x), _dp(u), float(dt), _ip(c), _dp(xn))
    return xn


def model_dynamics_partial(kind, x, u, dt, c):
    x = np.ascontiguousarray(x, np.float64); u = np.ascontiguousarray(u, np.float64)
    c = np.ascontiguousarray(c, np.int32); A = np.zeros(576); B = np.zeros(576)
    lib().orc_model_dynamics_partial(kind, _dp(x), _dp(u), float(dt), _ip(c), _dp(A), _dp(B))
    return A.reshape(24, 24).T.copy(), B.reshape(24, 24).T.copy()  # column-major -> [row, col]


def model_foot_position(kind, pos, eul, q, leg):
    pos = np.ascontiguousarray(pos, np.float64); eul = np.ascontiguousarray(eul, np.float64); q = np.ascontiguousarray(q, np.float64)
    p = np.zeros(3)
    lib().orc_model_foot_position(kind, _dp(pos), _dp(eul), _dp(q), int(leg), _dp(p))
    return p


def model_foot_jacobian(kind, pos, eul, q, leg):
    pos = np.ascontiguousarray(pos, np.float64); eul = np.ascontiguousarray(eul, np.float64); q = np.ascontiguousarray(q, np.float64)
    J = np.zeros(54)
    lib().orc_model_foot_jacobian(kind, _dp(pos), _dp(eul), _dp(q), int(leg), _dp(J))
    return J.reshape(18, 3).T.copy()


def model_hkd_state(kind, eul, pos, qJ, c):
    eul = np.ascontiguousarray(eul, np.float64); pos = np.ascontiguousarray(pos, np.float64); qJ = np.ascontiguousarray(qJ, np.float64)
    c = np.ascontiguousarray(c, np.int32); qd = np.zeros(12)
    lib().orc_model_hkd_state(kind, _dp(eul), _dp(pos), _dp(qJ), _ip(c), _dp(qd))
    return qd


class GaitTable:
    def __init__(self, npz):
        d = np.load(npz) if isinstance(npz, str) else npz
        self.dt = float(d["dt"])
        self.body_state = np.ascontiguousarray(d["body_state"], np.float32)
        self.qJ = np.ascontiguousarray(d["qJ"], np.float32)
        self.foot_placements = np.ascontiguousarray(d["foot_placements"], np.float32)
        self.grf = np.ascontiguousarray(d["grf"], np.float32)
        self.contact = np.ascontiguousarray(d["contact"], np.int32)
        self.n = self.body_state.shape[0]
        self.handle = lib().orc_table_create(self.n, C.c_float(self.dt), _fp(self.body_state), _fp(self.qJ),
                                             _fp(self.foot_placements), _fp(self.grf), _ip(self.contact))

    def __del__(self):
        try:
            lib().orc_table_destroy(self.handle)
        except Exception:
            pass


class Problem:
    """One HKD multi-phase problem (HKDProblem::initialization + MultiPhaseDDP)."""

    def __init__(self, table, k0=0, plan=0.6, model=None, cparams=None):
        self.table = table
        self.model = default_model() if model is None else model
        self._cp = cparams_array(**(cparams or {}))
        self.h = lib().orc_problem_create(table.handle, int(k0), C.c_float(plan), self.model, _dp(self._cp))
        self._refresh()

    def _refresh(self):
        self.n_phases = lib().orc_problem_n_phases(self.h)
        self.n_stages = lib().orc_problem_n_stages(self.h)
        self.phases = []
        for i in range(self.n_phases):
            hz = C.c_int(); c = (C.c_int * 4)(); cn = (C.c_int * 4)(); st = C.c_float(); ntd = C.c_int(); npth = C.c_int()
            lib().orc_problem_phase_info(self.h, i, C.byref(hz), c, cn, C.byref(st), C.byref(ntd), C.byref(npth))
            ss = C.c_int(); ht = C.c_int(); re = C.c_int(); no = C.c_int()
            lib().orc_problem_phase_flags(self.h, i, C.byref(ss), C.byref(ht), C.byref(re), C.byref(no))
            self.phases.append(dict(horizon=hz.value, contact=list(c), next_contact=list(cn), start_time=st.value,
                                    n_td=ntd.value, n_path=npth.value, ss_size=ss.value, has_tconstr=bool(ht.value),
                                    reach_end=bool(re.value), n_td_objects=no.value))
        self.n_states = self.n_stages + self.n_phases
        self.window_start = lib().orc_problem_window_start(self.h)

    def mpc_update(self):
        """HKDProblem::update (HKDProblem.cpp:117-222): shift the horizon by one MPC step."""
        lib().orc_mpc_update(self.h)
        self._refresh()

    def __del__(self):
        try:
            lib().orc_problem_destroy(self.h)
        except Exception:
            pass

    @property
    def x0(self):
        x = np.zeros(24); lib().orc_problem_get_x0(self.h, _dp(x)); return x

    @x0.setter
    def x0(self, v):
        v = np.ascontiguousarray(v, np.float64); lib().orc_problem_set_x0(self.h, _dp(v))

    def stage_reference(self, phase, k):
        xr = np.zeros(24); ur = np.zeros(24); br = np.zeros(12); fr = np.zeros(12); idx = C.c_int()
        lib().orc_problem_stage_reference(self.h, phase, k, _dp(xr), _dp(ur), _dp(br), _dp(fr), C.byref(idx))
        return xr, ur, br, fr, idx.value

    _SHAPES = {0: "s", 1: "s", 2: "s", 3: "s", 4: "s", 5: "s", 10: "u", 11: "u", 12: "u", 20: "m", 21: "m", 22: "m",
               24: "m", 25: "m", 26: "m", 27: "ms", 30: "u", 31: "u", 32: "l", 40: "pv", 41: "pm", 42: "p", 50: "p4", 51: "p8", 52: "g", 53: "r", 54: "p16"}
    NAMES = dict(Xbar=0, X=1, Xsim=2, Defect=3, dX=4, G=5, Ubar=10, U=11, dU=12, K=20, A=21, B=22, lxx=24, luu=25, lux=26,
                 H=27, lx=30, lu=31, l=32, Phix=40, Phixx=41, Phi=42, h=50, al=51, g=52, reb=53, td_by_leg=54)

    def get(self, name):
        which = self.NAMES[name]
        kind = self._SHAPES[which]
        N, S, P = self.n_stages, self.n_states, self.n_phases
        shape = {"s": (S, 24), "u": (N, 24), "m": (N, 24, 24), "ms": (S, 24, 24), "l": (N,), "pv": (P, 24), "pm": (P, 24, 24),
                 "p": (P,), "p4": (P, 4), "p8": (P, 4, 2), "g": (N, 20), "r": (N, 20, 2), "p16": (P, 4, 4)}[kind]
        out = np.zeros(shape)
        lib().orc_problem_get(self.h, which, _dp(out))
        if kind in ("m", "ms", "pm"):
            out = np.ascontiguousarray(np.swapaxes(out, -1, -2))  # column-major blocks -> [row, col]
        return out

    def set(self, name, arr):
        arr = np.ascontiguousarray(arr, np.float64)
        lib().orc_problem_set(self.h, self.NAMES[name], _dp(arr))

    def scalars(self):
        s = np.zeros(8); lib().orc_problem_scalars(self.h, _dp(s))
        return dict(zip(["actual_cost", "merit", "feas", "dV_1", "dV_2", "max_tconstr", "max_pconstr", "merit_rho"], s))

    # step-level API
    def hybrid_rollout(self, eps, opts=None):
        o = options_array(**(opts or {})); return bool(lib().orc_hybrid_rollout(self.h, float(eps), _dp(o)))

    def compute_cost(self, opts=None):
        o = options_array(**(opts or {})); lib().orc_compute_cost(self.h, _dp(o))

    def lq_approximation(self, opts=None):
        o = options_array(**(opts or {})); lib().orc_lq_approximation(self.h, _dp(o))

    def backward_sweep(self, reg):
        return bool(lib().orc_backward_sweep(self.h, float(reg)))

    def linear_rollout(self, eps, opts=None):
        o = options_array(**(opts or {})); lib().orc_linear_rollout(self.h, float(eps), _dp(o))

    def update_nominal(self):
        lib().orc_update_nominal(self.h)

    def solve(self, opts=None):
        o = options_array(**(opts or {}))
        summary = np.zeros(10); cap = 256; trace = np.zeros((cap, 16))
        n = lib().orc_solve(self.h, _dp(o), _dp(summary), _dp(trace), cap)
        return dict(zip(SUMMARY_COLS, summary)), trace[:n].copy()


def batch_solve(tables, k0, x0, plan=0.6, model=None, opts=None, cparams=None, n_threads=0, want_summaries=True):
    """One problem per std::thread.  tables: list of GaitTable (one per problem)."""
    n = len(tables)
    tp = (C.c_void_p * n)(*[t.handle for t in tables])
    k0 = np.ascontiguousarray(k0, np.int32)
    x0 = np.ascontiguousarray(x0, np.float64)
    o = options_array(**(opts or {})); cp = cparams_array(**(cparams or {}))
    summ = np.zeros((n, 10)) if want_summaries else None
    model = default_model() if model is None else model
    wall = lib().orc_batch_solve(tp, _ip(k0), _dp(x0), n, C.c_float(plan), model, _dp(o), _dp(cp), int(n_threads),
                                 _dp(summ) if want_summaries else None)
    return wall, summ


def batch_solve_traj(tables, k0, x0, max_nodes, max_stages, k_rows=8, plan=0.6, model=None, opts=None, cparams=None, n_threads=0, variant=None):
    """batch_solve that also returns the solutions: dict(wall, summary [n,10], Xbar [n,max_nodes,24], Ubar [n,max_stages,24],
    K [n,k_rows,24,24] (row, col), n_trials [n]).  Rows beyond a problem's own node / stage count are zero."""
    n = len(tables)
    tp = (C.c_void_p * n)(*[t.handle for t in tables])
    k0 = np.ascontiguousarray(k0, np.int32)
    x0 = np.ascontiguousarray(x0, np.float64)
    o = options_array(**(opts or {})); cp = cparams_array(**(cparams or {}))
    summ = np.zeros((n, 10)); Xb = np.zeros((n, max_nodes, 24)); Ub = np.zeros((n, max_stages, 24))
    K = np.zeros((n, k_rows, 24, 24)) if k_rows > 0 else None
    ntr = np.zeros(n, np.int32)
    model = default_model() if model is None else model
    wall = _variant_lib(variant).orc_batch_solve_traj(tp, _ip(k0), _dp(x0), n, C.c_float(plan), model, _dp(o), _dp(cp), int(n_threads), _dp(summ),
                                      _dp(Xb), _dp(Ub), int(max_nodes), int(max_stages), _dp(K) if K is not None else None, int(k_rows), _ip(ntr))
    if K is not None:
        K = np.ascontiguousarray(np.swapaxes(K, -1, -2))
    return dict(wall=wall, summary=summ, Xbar=Xb, Ubar=Ub, K=K, n_trials=ntr)



# ---- generic SinglePhase<T, xs, us, ys> sweeps on plug-in outputs (oracle/single_phase_generic.hpp) ----
GENERIC_INPUTS = ["A", "B", "C", "D", "lx", "lu", "ly", "lxx", "luu", "lux", "lyy", "Phix", "Phixx", "Defect"]


def generic_shapes(xs, us, ys, N):
    """[row, col] shapes of one phase's plug-in outputs (matrices are handed to C column-major)."""
    return dict(A=(N, xs, xs), B=(N, xs, us), C=(N, ys, xs), D=(N, ys, us), lx=(N, xs), lu=(N, us), ly=(N, ys), lxx=(N, xs, xs),
                luu=(N, us, us), lux=(N, us, xs), lyy=(N, ys, ys), Phix=(xs,), Phixx=(xs, xs), Defect=(N + 1, xs))


def _colmajor(a):
    a = np.asarray(a, np.float64)
    return np.ascontiguousarray(np.swapaxes(a, -1, -2)) if a.ndim >= 2 and a.shape[-1] != 0 else np.ascontiguousarray(a)


def _generic_pack(xs, us, ys, N, data):
    keep, ptrs = [], (C.POINTER(C.c_double) * len(GENERIC_INPUTS))()
    shp = generic_shapes(xs, us, ys, N)
    for i, nm in enumerate(GENERIC_INPUTS):
        a = np.asarray(data[nm], np.float64) if nm in data else np.zeros(shp[nm])
        assert a.shape == shp[nm], (nm, a.shape, shp[nm])
        a = _colmajor(a) if nm in ("A", "B", "C", "D", "lxx", "luu", "lux", "lyy", "Phixx") else np.ascontiguousarray(a)
        keep.append(a)
        ptrs[i] = _dp(a)
    return keep, ptrs


def generic_backward_sweep(xs, us, ys, N, data, reg, Gprime=None, Hprime=None):
    """SinglePhase<double, xs, us, ys>::backward_sweep -> dict(success, dU, K [N, us, xs], G, H [N+1, xs, xs], dV_1, dV_2)."""
    L = lib()
    L.orc_generic_backward_sweep.argtypes = [C.c_int] * 4 + [C.POINTER(C.POINTER(C.c_double)), C.c_double] + [C.POINTER(C.c_double)] * 7
    keep, ptrs = _generic_pack(xs, us, ys, N, data)
    Gp = np.zeros(xs) if Gprime is None else np.ascontiguousarray(Gprime, np.float64)
    Hp = np.zeros((xs, xs)) if Hprime is None else _colmajor(Hprime)
    dU, K, G, H, dV = np.zeros((N, us)), np.zeros((N, xs, us)), np.zeros((N + 1, xs)), np.zeros((N + 1, xs, xs)), np.zeros(2)
    ok = L.orc_generic_backward_sweep(xs, us, ys, N, ptrs, float(reg), _dp(Gp), _dp(Hp), _dp(dU), _dp(K), _dp(G), _dp(H), _dp(dV))
    return dict(success=bool(ok), dU=dU, K=np.ascontiguousarray(np.swapaxes(K, 1, 2)), G=G,
                H=np.ascontiguousarray(np.swapaxes(H, 1, 2)), dV_1=dV[0], dV_2=dV[1])


def generic_linear_rollout(xs, us, ys, N, data, eps, dU, K, dx_init=None):
    """SinglePhase<double, xs, us, ys>::linear_rollout -> dict(dX, dV_1, dV_2)."""
    L = lib()
    L.orc_generic_linear_rollout.argtypes = [C.c_int] * 4 + [C.POINTER(C.POINTER(C.c_double)), C.c_double] + [C.POINTER(C.c_double)] * 5
    keep, ptrs = _generic_pack(xs, us, ys, N, data)
    dx0 = np.zeros(xs) if dx_init is None else np.ascontiguousarray(dx_init, np.float64)
    dUc, Kc = np.ascontiguousarray(dU, np.float64), _colmajor(K)
    dX, dV = np.zeros((N + 1, xs)), np.zeros(2)
    L.orc_generic_linear_rollout(xs, us, ys, N, ptrs, float(eps), _dp(dx0), _dp(dUc), _dp(Kc), _dp(dX), _dp(dV))
    return dict(dX=dX, dV_1=dV[0], dV_2=dV[1])


def hardware_concurrency():
    return lib().orc_hardware_concurrency()


def mpc_command(P, n_steps=8):
    """HKDMPCSolver::publish_mpc_cmd + update_foot_placement (HKDMPC/HKDMPC.cpp:207-298) restated on the oracle's
    trajectories: returns dict(hkd_controls [n,24], des_body_state [n,12], feedback [n,12,12], contacts [n,4],
    foot_placement [12], foot_found [4]) in float32 like the LCM message."""
    Xbar, Ubar, K = P.get("Xbar"), P.get("Ubar"), P.get("K")  # K: [stages, 24, 24] with K[s][i, j] = dense gain
    node_off, stage_off, no, so = [], [], 0, 0
    for ph in P.phases:
        node_off.append(no); stage_off.append(so)
        no += ph["horizon"] + 1; so += ph["horizon"]
    out = dict(hkd_controls=np.zeros((n_steps, 24), np.float32), des_body_state=np.zeros((n_steps, 12), np.float32),
               feedback=np.zeros((n_steps, 12, 12), np.float32), contacts=np.zeros((n_steps, 4), np.int32),
               foot_placement=np.zeros(12, np.float32), foot_found=np.zeros(4, np.int32))
    k = s = i = 0
    while k < n_steps:
        if s >= P.phases[i]["horizon"]:
            s = 0; i += 1
        out["hkd_controls"][k] = Ubar[stage_off[i] + s]
        out["des_body_state"][k] = Xbar[node_off[i] + s][:12]
        out["feedback"][k] = K[stage_off[i] + s][:12, :12]
        out["contacts"][k] = P.phases[i]["contact"]
        s += 1; k += 1
    for i in range(P.n_phases - 1):
        c, cn = P.phases[i]["contact"], P.phases[i + 1]["contact"]
        for l in range(4):
            if not out["foot_found"][l] and c[l] == 0 and cn[l] == 1:
                out["foot_placement"][3 * l:3 * l + 3] = Xbar[node_off[i + 1]][12 + 3 * l:15 + 3 * l]
                out["foot_found"][l] = 1
        if i >= 4:
            break
    return out
