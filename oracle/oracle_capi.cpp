// ORACLE (test infrastructure only — never linked by the product).
// Plain C entry points so tests/ and bench.py's cpu_baseline leg can drive the
// oracle through ctypes.  Layouts of the flat trajectory getters:
//   states   : phases concatenated, each phase contributes horizon+1 rows of 24
//   controls : phases concatenated, each phase contributes horizon   rows of 24
//   matrices : 576 doubles, column-major (Eigen default), one per control stage
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>
#include "hsddp_oracle.hpp"
#include "single_phase_generic.hpp"

using namespace oracle;

namespace {
void options_from_array(const double* o, Options& opt) {
    // order documented in tests/oracle_py.py OPTION_FIELDS
    opt.alpha = o[0]; opt.gamma = o[1]; opt.update_penalty = o[2]; opt.update_relax = o[3];
    opt.update_regularization = o[4]; opt.update_ReB = o[5]; opt.max_DDP_iter = (int)o[6]; opt.max_AL_iter = (int)o[7];
    opt.cost_thresh = o[8]; opt.tconstr_thresh = o[9]; opt.pconstr_thresh = o[10]; opt.dynamics_feas_thresh = o[11];
    opt.merit_scale = o[12]; opt.merit_offset = o[13]; opt.AL_active = (int)o[14]; opt.ReB_active = (int)o[15]; opt.MS = (int)o[16];
}
void cparams_from_array(const double* c, ConstraintParams& cp) {
    cp.grf_delta = c[0]; cp.grf_delta_min = c[1]; cp.grf_eps = c[2];
    cp.td_sigma = c[3]; cp.td_sigma_max = c[4]; cp.td_lambda = c[5]; cp.mu = c[6];
}
}  // namespace

extern "C" {

int orc_load_ref(const char* path) { return casadi_ref().load(path) ? 1 : 0; }
int orc_ref_loaded() { return casadi_ref().handle != nullptr; }

// ---- model level ----
void orc_model_dynamics(int kind, const double* x, const double* u, double dt, const int* c, double* xn) {
    Model m; m.kind = (ModelKind)kind; m.dynamics(x, u, dt, c, xn);
}
void orc_model_dynamics_partial(int kind, const double* x, const double* u, double dt, const int* c, double* A, double* B) {
    Model m; m.kind = (ModelKind)kind; m.dynamics_partial(x, u, dt, c, A, B);
}
void orc_model_foot_position(int kind, const double* pos, const double* eul, const double* q, int leg, double* p) {
    Model m; m.kind = (ModelKind)kind; m.foot_position(pos, eul, q, leg, p);
}
void orc_model_foot_jacobian(int kind, const double* pos, const double* eul, const double* q, int leg, double* J) {
    Model m; m.kind = (ModelKind)kind; m.foot_jacobian(pos, eul, q, leg, J);
}
void orc_model_hkd_state(int kind, const double* eul, const double* pos, const double* qJ, const int* c, double* qd) {
    Model m; m.kind = (ModelKind)kind; m.hkd_state(eul, pos, qJ, c, qd);
}

// ---- linear algebra pieces (unit tests) ----
int orc_ldlt_is_positive(const double* M) { Mat24 m; std::memcpy(m.m, M, sizeof m.m); return ldlt_is_positive(m) ? 1 : 0; }
void orc_inverse(const double* M, double* Inv) { Mat24 m, r; std::memcpy(m.m, M, sizeof m.m); inverse_partial_piv_lu(m, r); std::memcpy(Inv, r.m, sizeof r.m); }

// ---- gait table ----
void* orc_table_create(int n, float dt, const float* body, const float* qJ, const float* foot, const float* grf, const int* contact) {
    GaitTable* t = new GaitTable();
    t->n = n; t->dt = dt;
    t->body_state.assign(body, body + 12 * (size_t)n);
    t->qJ.assign(qJ, qJ + 12 * (size_t)n);
    t->foot_placements.assign(foot, foot + 12 * (size_t)n);
    t->grf.assign(grf, grf + 12 * (size_t)n);
    t->contact.assign(contact, contact + 4 * (size_t)n);
    return t;
}
void orc_table_destroy(void* t) { delete (GaitTable*)t; }

// ---- problem ----
void* orc_problem_create(void* table, int k0, float plan, int model_kind, const double* cparams) {
    Problem* p = new Problem();
    ConstraintParams cp;
    if (cparams) cparams_from_array(cparams, cp);
    p->build((GaitTable*)table, k0, plan, (ModelKind)model_kind, cp);
    return p;
}
void orc_problem_destroy(void* p) { delete (Problem*)p; }
int orc_problem_n_phases(void* p) { return (int)((Problem*)p)->phases.size(); }
int orc_problem_n_stages(void* p) { int n = 0; for (auto& ph : ((Problem*)p)->phases) n += ph.horizon; return n; }
void orc_problem_phase_info(void* p, int i, int* horizon, int* contact, int* next_contact, float* start_time, int* n_td, int* n_path) {
    const Phase& ph = ((Problem*)p)->phases[i];
    *horizon = ph.horizon; *start_time = ph.start_time; *n_td = ph.n_td_total(); *n_path = ph.n_path;
    for (int l = 0; l < 4; ++l) { contact[l] = ph.contact[l]; next_contact[l] = ph.next_contact[l]; }
}
void orc_problem_set_x0(void* p, const double* x0) { std::memcpy(((Problem*)p)->x0.v, x0, 24 * sizeof(double)); }
void orc_problem_get_x0(void* p, double* x0) { std::memcpy(x0, ((Problem*)p)->x0.v, 24 * sizeof(double)); }
// per-stage reference rows exactly as the cost callbacks see them (running stages then the terminal stage of each phase)
void orc_problem_stage_reference(void* pv, int phase, int k, double* xr, double* ur, double* body_r, double* foot_r, int* idx_out) {
    Problem* p = (Problem*)pv;
    const Phase& ph = p->phases[phase];
    const float t = (float)((double)ph.t_offset + k * ph.dt);
    int idx;
    p->reference_at_t(t, xr, ur, &idx);
    std::memcpy(body_r, p->ref.body_state(idx), 12 * sizeof(double));
    std::memcpy(foot_r, p->ref.foot(idx), 12 * sizeof(double));
    *idx_out = idx;
}

// which: 0 Xbar 1 X 2 Xsim 3 Defect 4 dX 5 G (state-shaped) ; 10 Ubar 11 U 12 dU (control-shaped)
//        20 K 21 A 22 B 23 H(k) 24 lxx 25 luu 26 lux (one 576-block per control stage; H uses states layout via which=27)
void orc_problem_get(void* pv, int which, double* out) {
    Problem* p = (Problem*)pv;
    size_t o = 0;
    for (auto& ph : p->phases) {
        const int N = ph.horizon;
        auto put_states = [&](const std::vector<Vec24>& v) { for (int k = 0; k <= N; ++k) { std::memcpy(out + o, v[k].v, 192); o += 24; } };
        auto put_ctrls = [&](const std::vector<Vec24>& v) { for (int k = 0; k < N; ++k) { std::memcpy(out + o, v[k].v, 192); o += 24; } };
        switch (which) {
            case 0: put_states(ph.Xbar); break;
            case 1: put_states(ph.X); break;
            case 2: put_states(ph.Xsim); break;
            case 3: put_states(ph.Defect); break;
            case 4: put_states(ph.dX); break;
            case 5: put_states(ph.G); break;
            case 10: put_ctrls(ph.Ubar); break;
            case 11: put_ctrls(ph.U); break;
            case 12: put_ctrls(ph.dU); break;
            case 20: for (int k = 0; k < N; ++k) { std::memcpy(out + o, ph.K[k].m, 4608); o += 576; } break;
            case 21: for (int k = 0; k < N; ++k) { std::memcpy(out + o, ph.A[k].m, 4608); o += 576; } break;
            case 22: for (int k = 0; k < N; ++k) { std::memcpy(out + o, ph.B[k].m, 4608); o += 576; } break;
            case 24: for (int k = 0; k < N; ++k) { std::memcpy(out + o, ph.rcost[k].lxx.m, 4608); o += 576; } break;
            case 25: for (int k = 0; k < N; ++k) { std::memcpy(out + o, ph.rcost[k].luu.m, 4608); o += 576; } break;
            case 26: for (int k = 0; k < N; ++k) { std::memcpy(out + o, ph.rcost[k].lux.m, 4608); o += 576; } break;
            case 27: for (int k = 0; k <= N; ++k) { std::memcpy(out + o, ph.H[k].m, 4608); o += 576; } break;
            case 30: for (int k = 0; k < N; ++k) { std::memcpy(out + o, ph.rcost[k].lx.v, 192); o += 24; } break;
            case 31: for (int k = 0; k < N; ++k) { std::memcpy(out + o, ph.rcost[k].lu.v, 192); o += 24; } break;
            case 32: for (int k = 0; k < N; ++k) { out[o++] = ph.rcost[k].l; } break;
            case 40: std::memcpy(out + o, ph.tcost.Phix.v, 192); o += 24; break;
            case 41: std::memcpy(out + o, ph.tcost.Phixx.m, 4608); o += 576; break;
            case 42: out[o++] = ph.tcost.Phi; break;
            case 50: {  // touchdown values of the FIRST constraint object, packed by constraint
                for (int i = 0; i < 4; ++i) out[o++] = (!ph.tds.empty() && i < ph.tds[0].n_td) ? ph.tds[0].h[i] : 0.0;
            } break;
            case 51: {  // (sigma, lambda) of the first constraint object (initial values when the phase has none)
                for (int i = 0; i < 4; ++i) {
                    out[o++] = ph.tds.empty() ? p->cparams.td_sigma : ph.tds[0].al[i].sigma;
                    out[o++] = ph.tds.empty() ? p->cparams.td_lambda : ph.tds[0].al[i].lambda;
                }
            } break;
            case 54: {  // per LEG: [sum over constraint objects of sigma, sum of lambda, h, number of objects on the leg]
                for (int l = 0; l < 4; ++l) {
                    double ss = 0, sl = 0, hh = 0; int cnt = 0;
                    for (auto& td : ph.tds)
                        for (int i = 0; i < td.n_td; ++i)
                            if (td.td_legs[i] == l) { ss += td.al[i].sigma; sl += td.al[i].lambda; hh = td.h[i]; ++cnt; }
                    out[o++] = ss; out[o++] = sl; out[o++] = hh; out[o++] = cnt;
                }
            } break;
            case 52: for (int k = 0; k < N; ++k) for (int i = 0; i < 20; ++i) out[o++] = (i < ph.n_path) ? ph.g[(size_t)k * ph.n_path + i] : 0.0; break;
            case 53:  // ReB parameters (eps, delta) per stage, 5 rows per LEG like the GPU layout; swing legs keep the initial values
                for (int k = 0; k < N; ++k) {
                    for (int i = 0; i < 20; ++i) { out[o + 2 * i] = p->cparams.grf_eps; out[o + 2 * i + 1] = p->cparams.grf_delta; }
                    for (int sl = 0; sl < ph.n_stance; ++sl)
                        for (int r = 0; r < 5; ++r) {
                            const RebParam& q = ph.reb[(size_t)k * ph.n_path + 5 * sl + r];
                            out[o + 2 * (5 * ph.stance_legs[sl] + r)] = q.eps; out[o + 2 * (5 * ph.stance_legs[sl] + r) + 1] = q.delta;
                        }
                    o += 40;
                }
                break;
            default: break;
        }
    }
}
// overwrite Xbar/X (which=0/1) or Ubar/U (10/11) — used to seed warm starts in tests
void orc_problem_set(void* pv, int which, const double* in) {
    Problem* p = (Problem*)pv;
    size_t o = 0;
    for (auto& ph : p->phases) {
        const int N = ph.horizon;
        if (which == 0) for (int k = 0; k <= N; ++k) { std::memcpy(ph.Xbar[k].v, in + o, 192); o += 24; }
        if (which == 1) for (int k = 0; k <= N; ++k) { std::memcpy(ph.X[k].v, in + o, 192); o += 24; }
        if (which == 10) for (int k = 0; k < N; ++k) { std::memcpy(ph.Ubar[k].v, in + o, 192); o += 24; }
        if (which == 11) for (int k = 0; k < N; ++k) { std::memcpy(ph.U[k].v, in + o, 192); o += 24; }
    }
}

// scalars: [actual_cost, merit, feas, dV_1, dV_2, max_tconstr, max_pconstr, merit_rho]
void orc_problem_scalars(void* pv, double* s) {
    Problem* p = (Problem*)pv;
    s[0] = p->actual_cost; s[1] = p->merit; s[2] = p->feas; s[3] = p->dV_1; s[4] = p->dV_2;
    s[5] = p->max_tconstr; s[6] = p->max_pconstr; s[7] = p->merit_rho;
}

// ---- step-level API (MultiPhaseDDP public methods) ----
int orc_hybrid_rollout(void* p, double eps, const double* o) { Options opt; options_from_array(o, opt); return ((Problem*)p)->hybrid_rollout(eps, opt) ? 1 : 0; }
void orc_compute_cost(void* p, const double* o) { Options opt; options_from_array(o, opt); Problem* q = (Problem*)p; q->compute_cost(opt); q->feas = q->measure_dynamics_feasibility(); }
void orc_lq_approximation(void* p, const double* o) { Options opt; options_from_array(o, opt); ((Problem*)p)->LQ_approximation(opt); }
int orc_backward_sweep(void* p, double reg) { return ((Problem*)p)->backward_sweep(reg) ? 1 : 0; }
void orc_linear_rollout(void* p, double eps, const double* o) { Options opt; options_from_array(o, opt); ((Problem*)p)->linear_rollout(eps, opt); }
void orc_update_nominal(void* p) { ((Problem*)p)->update_nominal_trajectory(); }
// HKDProblem::update: receding-horizon shift by one MPC step
void orc_mpc_update(void* p) { ((Problem*)p)->update(); }
int orc_problem_phase_flags(void* p, int i, int* ss_size, int* has_tconstr, int* reach_end, int* n_td_objects) {
    const Phase& ph = ((Problem*)p)->phases[i];
    *ss_size = ph.ss_size; *has_tconstr = ph.has_tconstr ? 1 : 0; *reach_end = ph.reach_end ? 1 : 0; *n_td_objects = (int)ph.tds.size();
    return 0;
}
int orc_problem_window_start(void* p) { return ((Problem*)p)->ref.k0; }

// ---- full solve ----
// summary: [status, n_iter, n_outer, n_sweeps, cost, feas, max_tconstr, max_pconstr, cost0, feas0]
// trace  : n_iter rows x 16: outer, inner, cost_before, feas_before, reg_after, n_sweeps, dV_1, dV_2, merit_rho,
//          eps_accepted, n_trials, cost_after, feas_after, max_tconstr, max_pconstr, 0
int orc_solve(void* pv, const double* o, double* summary, double* trace, int trace_cap) {
    Options opt; options_from_array(o, opt);
    SolveResult r;
    ((Problem*)pv)->solve(opt, r);
    summary[0] = r.status; summary[1] = r.n_iter; summary[2] = r.n_outer; summary[3] = r.n_sweeps;
    summary[4] = r.cost; summary[5] = r.feas; summary[6] = r.max_tconstr; summary[7] = r.max_pconstr;
    summary[8] = r.cost0; summary[9] = r.feas0;
    int n = (int)r.trace.size();
    for (int i = 0; i < n && i < trace_cap; ++i) {
        const IterRecord& t = r.trace[i];
        double* row = trace + 16 * (size_t)i;
        row[0] = t.outer; row[1] = t.inner; row[2] = t.cost_before; row[3] = t.feas_before; row[4] = t.reg_used; row[5] = t.n_sweeps;
        row[6] = t.dV_1; row[7] = t.dV_2; row[8] = t.merit_rho; row[9] = t.eps_accepted; row[10] = t.n_trials;
        row[11] = t.cost_after; row[12] = t.feas_after; row[13] = t.max_tconstr; row[14] = t.max_pconstr; row[15] = 0;
    }
    return n;
}

// ---- CPU baseline: one problem per std::thread ----
// tables: n_prob pointers (one gait table per problem), k0: window starts, x0: n_prob x 24.
// Returns wall seconds; per-problem summary rows (10 doubles) written to `summaries` when non-null.
// Optional trajectory outputs for parity checks at scale (bench.py, tests): xbar_out [n_prob][max_nodes][24],
// ubar_out [n_prob][max_stages][24] (rows beyond a problem's own count are left untouched), k_out [n_prob][k_rows][576]
// = the column-major gains of the first k_rows stages (what HKDMPC.cpp:245-275 ships), trials_out [n_prob] = total
// line-search trials.
double orc_batch_solve_traj(void** tables, const int* k0, const double* x0, int n_prob, float plan, int model_kind,
                            const double* o, const double* cparams, int n_threads, double* summaries,
                            double* xbar_out, double* ubar_out, int max_nodes, int max_stages, double* k_out, int k_rows, int* trials_out) {
    Options opt; options_from_array(o, opt);
    ConstraintParams cp;
    if (cparams) cparams_from_array(cparams, cp);
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    std::atomic<int> next(0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) {
        pool.emplace_back([&]() {
            for (;;) {
                int i = next.fetch_add(1);
                if (i >= n_prob) break;
                Problem p;
                p.build((GaitTable*)tables[i], k0[i], plan, (ModelKind)model_kind, cp);
                std::memcpy(p.x0.v, x0 + 24 * (size_t)i, 24 * sizeof(double));
                SolveResult r;
                p.solve(opt, r);
                if (summaries) {
                    double* s = summaries + 10 * (size_t)i;
                    s[0] = r.status; s[1] = r.n_iter; s[2] = r.n_outer; s[3] = r.n_sweeps; s[4] = r.cost; s[5] = r.feas;
                    s[6] = r.max_tconstr; s[7] = r.max_pconstr; s[8] = r.cost0; s[9] = r.feas0;
                }
                if (trials_out) { int nt = 0; for (auto& rec : r.trace) nt += rec.n_trials; trials_out[i] = nt; }
                if (xbar_out || ubar_out || k_out) {
                    size_t node = 0, stage = 0;
                    for (auto& ph : p.phases) {
                        for (int k = 0; k <= ph.horizon; ++k, ++node)
                            if (xbar_out && (int)node < max_nodes) std::memcpy(xbar_out + ((size_t)i * max_nodes + node) * 24, ph.Xbar[k].v, 192);
                        for (int k = 0; k < ph.horizon; ++k, ++stage) {
                            if (ubar_out && (int)stage < max_stages) std::memcpy(ubar_out + ((size_t)i * max_stages + stage) * 24, ph.Ubar[k].v, 192);
                            if (k_out && (int)stage < k_rows) std::memcpy(k_out + ((size_t)i * k_rows + stage) * 576, ph.K[k].m, 4608);
                        }
                    }
                }
            }
        });
    }
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

double orc_batch_solve(void** tables, const int* k0, const double* x0, int n_prob, float plan, int model_kind,
                       const double* o, const double* cparams, int n_threads, double* summaries) {
    return orc_batch_solve_traj(tables, k0, x0, n_prob, plan, model_kind, o, cparams, n_threads, summaries, nullptr, nullptr, 0, 0, nullptr, 0, nullptr);
}

// ---- generic SinglePhase<T, xs, us, ys> sweeps on plug-in outputs (single_phase_generic.hpp) ----
// in[14] = {A, B, C, D, lx, lu, ly, lxx, luu, lux, lyy, Phix, Phixx, Defect}; C, D, ly, lyy may be null when ys == 0
static generic::PhaseData generic_phase(int xs, int us, int ys, int N, const double* const* in) {
    generic::PhaseData p;
    p.xs = xs; p.us = us; p.ys = ys; p.N = N;
    p.A = in[0]; p.B = in[1]; p.C = in[2]; p.D = in[3]; p.lx = in[4]; p.lu = in[5]; p.ly = in[6]; p.lxx = in[7]; p.luu = in[8];
    p.lux = in[9]; p.lyy = in[10]; p.Phix = in[11]; p.Phixx = in[12]; p.Defect = in[13];
    return p;
}
int orc_generic_backward_sweep(int xs, int us, int ys, int N, const double* const* in, double reg, const double* Gprime,
                               const double* Hprime, double* dU, double* K, double* G, double* H, double* dV) {
    return generic::backward_sweep(generic_phase(xs, us, ys, N, in), reg, Gprime, Hprime, dU, K, G, H, dV) ? 1 : 0;
}
void orc_generic_linear_rollout(int xs, int us, int ys, int N, const double* const* in, double eps, const double* dx_init,
                                const double* dU, const double* K, double* dX, double* dV) {
    generic::linear_rollout(generic_phase(xs, us, ys, N, in), eps, dx_init, dU, K, dX, dV);
}

int orc_hardware_concurrency() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
