// ORACLE (test infrastructure only — never linked by the product).
// See hsddp_oracle.hpp for the list of reference files restated here.
#include "hsddp_oracle.hpp"
#include <algorithm>
#include <cmath>
#include <cstdio>

namespace oracle {

// ---------------------------------------------------------------------------
// HSDDP_Utils.h:46-78 — float-tolerance comparisons used by the phase split
// ---------------------------------------------------------------------------
static inline bool approx_eq_scalar(float n1, float n2) {
    float tol = 1e-6f;
    float err = std::fabs(n1 - n2);
    return err <= tol;
}
static inline bool approx_leq_scalar(float n1, float n2) { return n1 < n2 || approx_eq_scalar(n1, n2); }
static inline bool approx_geq_scalar(float n1, float n2) { return n1 > n2 || approx_eq_scalar(n1, n2); }

// ---------------------------------------------------------------------------
// QuadReference (Reference/QuadReference.cpp:6-26,65-100)
// ---------------------------------------------------------------------------
void QuadReference::initialize(const GaitTable* table, int window_start, float plan_horizon) {
    tp = table;
    k0 = window_start;
    dt = table->dt;
    sz = (int)std::round(plan_horizon / dt) + 1;
    t_cur = 0.f;
    dur = plan_horizon;
    start_time = t_cur;
    end_time = t_cur + dur;
}

void QuadReference::step(float dt_sim) {
    for (int i = 1; approx_leq_scalar((float)i * dt, dt_sim); i++) {
        k0++;
        t_cur += dt;
        start_time = t_cur;
        end_time = t_cur + dur;
    }
}

int QuadReference::index_at_t(float t) const {
    int k = (int)std::floor(t / dt);
    float rem = t - (float)k * dt;
    if ((double)rem > 0.5 * (double)dt) k++;
    if (k > sz) k = sz;  // "queried reference out of scope": clamp to the extra sample
    return k;
}

const int* QuadReference::contact_at_t(float t) const { return contact(index_at_t(t)); }

// ---------------------------------------------------------------------------
// HKDProblem::initialization (HKDProblem.cpp:15-111) + create_problem_one_phase
// (:225-265) + add_tconstr_one_phase (:268-310)
// ---------------------------------------------------------------------------
void Problem::build(const GaitTable* table, int window_start, float plan_dur, ModelKind kind, const ConstraintParams& cp) {
    model.kind = kind;
    cparams = cp;
    plan_duration = plan_dur;
    dt_sim = 0.01f;
    dt_mpc = dt_sim * 1;
    ref.initialize(table, window_start, plan_duration);
    phases.clear();

    int contact_prev[4], contact_cur[4];
    float phase_start_time = 0.f, phase_end_time = 0.f;
    float t = 0.f;
    std::copy(ref.contact_at_t(t), ref.contact_at_t(t) + 4, contact_prev);
    while (approx_leq_scalar(t, plan_duration)) {
        std::copy(ref.contact_at_t(t), ref.contact_at_t(t) + 4, contact_cur);
        bool change = false;
        for (int l = 0; l < 4; ++l) change = change || (contact_cur[l] != contact_prev[l]);
        if (change || approx_geq_scalar(t, plan_duration)) {
            phase_end_time = t;
            Phase ph;
            ph.horizon = (int)std::round((phase_end_time - phase_start_time) / dt_sim);
            ph.start_time = phase_start_time;
            ph.end_time = phase_end_time;
            ph.reach_end = false;  // (contact_prev != contact_prev).any(): always false, HKDProblem.cpp:59
            std::copy(contact_prev, contact_prev + 4, ph.contact);
            phases.push_back(ph);
            std::copy(contact_cur, contact_cur + 4, contact_prev);
            phase_start_time = phase_end_time;
        }
        t += dt_sim;
    }

    const int n_phases = (int)phases.size();
    for (int i = 0; i < n_phases; ++i) {
        Phase& ph = phases[i];
        create_phase(ph, ph.horizon);
        const int N = ph.horizon;
        // initial guess: state reference, zero control (HKDProblem.cpp:84-90)
        for (int k = 0; k <= N; ++k) {
            float tk = ph.start_time + (float)k * dt_sim;
            double xr[24], ur[24];
            reference_at_t(tk, xr, ur, nullptr);
            for (int j = 0; j < 24; ++j) { ph.X[k][j] = xr[j]; ph.Xbar[k][j] = xr[j]; }
        }
    }
    for (int i = 0; i < n_phases; ++i) {
        add_tconstr(i);  // every phase, the last one with the contact at plan_duration + dt_mpc (Q17)
        phases[i].t_offset = phases[i].start_time - phases[0].start_time;
        phases[i].ss_size = phases[i].horizon + 1;  // update_SS_config(horizon + 1), HKDProblem.cpp:104
    }
    default_x0(x0);
}

// Trajectory<T,24,24,0>::create_data (TrajectoryManagement.cpp:5-35) + create_problem_one_phase (HKDProblem.cpp:225-265)
// + SinglePhase::initialization (SinglePhase.cpp:22-35: the shooting set starts EMPTY)
void Problem::create_phase(Phase& ph, int horizon) {
    const int N = horizon;
    ph.horizon = N;
    ph.dt = (double)dt_sim;  // Trajectory(dt_sim, horizon): float widened to T
    ph.Xbar.assign(N + 1, Vec24{}); ph.X.assign(N + 1, Vec24{}); ph.Xsim.assign(N + 1, Vec24{});
    ph.Defect.assign(N + 1, Vec24{}); ph.Defect_bar.assign(N + 1, Vec24{}); ph.dX.assign(N + 1, Vec24{});
    ph.G.assign(N + 1, Vec24{});
    ph.Ubar.assign(N, Vec24{}); ph.U.assign(N, Vec24{}); ph.dU.assign(N, Vec24{});
    ph.A.assign(N + 1, Mat24{}); ph.B.assign(N, Mat24{}); ph.H.assign(N + 1, Mat24{}); ph.K.assign(N + 1, Mat24{});
    ph.rcost.assign(N, RCost{});
    for (auto& r : ph.rcost) r.zero();
    ph.tcost.zero();
    // GRF constraint on phases with a stance leg (HKDProblem.cpp:255-263)
    ph.n_stance = 0;
    for (int l = 0; l < 4; ++l) if (ph.contact[l] == 1) ph.stance_legs[ph.n_stance++] = l;
    ph.n_path = 5 * ph.n_stance;
    ph.g.assign((size_t)N * ph.n_path, 0.0);
    ph.reb.assign((size_t)N * ph.n_path, RebParam{cparams.grf_delta, cparams.grf_delta_min, cparams.grf_eps});
    ph.path_max_violation = 0.0;
    ph.tds.clear();
    ph.has_tconstr = false;
    ph.ss_size = 0;
    ph.x_init.zero(); ph.dx_init.zero();
}

// add_tconstr_one_phase (HKDProblem.cpp:268-310): binds the reset map to (contact, next contact) and ADDS a touchdown
// constraint object if a leg goes 0 -> 1.  Called for every phase at initialisation and again by update() when the last
// phase reaches its end (Q17).
void Problem::add_tconstr(int idx) {
    Phase& ph = phases[idx];
    const int n_phases = (int)phases.size();
    if (idx < n_phases - 1) std::copy(phases[idx + 1].contact, phases[idx + 1].contact + 4, ph.next_contact);
    else {
        const int* cn = ref.contact_at_t(plan_duration + dt_mpc);
        std::copy(cn, cn + 4, ph.next_contact);
    }
    ph.has_tconstr = true;
    Phase::TdSet td;
    for (int l = 0; l < 4; ++l)
        if (ph.contact[l] == 0 && ph.next_contact[l] == 1) td.td_legs[td.n_td++] = l;
    if (td.n_td > 0) {
        for (int c = 0; c < 4; ++c) {
            td.h[c] = 0; td.hx[c].zero();
            td.al[c] = AlParam{cparams.td_lambda, cparams.td_sigma, cparams.td_sigma_max};
        }
        td.max_violation = 0.0;
        ph.tds.push_back(td);
    }
}

// SinglePhase::pop_front: Trajectory::pop_front (TrajectoryManagement.cpp:118-146) + PathConstraintBase::pop_front
// (ConstraintsBase.h:271-275)
void Problem::phase_pop_front(Phase& ph) {
    ph.Xbar.erase(ph.Xbar.begin()); ph.X.erase(ph.X.begin()); ph.Xsim.erase(ph.Xsim.begin());
    ph.Defect.erase(ph.Defect.begin()); ph.Defect_bar.erase(ph.Defect_bar.begin()); ph.dX.erase(ph.dX.begin());
    ph.G.erase(ph.G.begin());
    ph.Ubar.erase(ph.Ubar.begin()); ph.U.erase(ph.U.begin()); ph.dU.erase(ph.dU.begin());
    ph.A.erase(ph.A.begin()); ph.B.erase(ph.B.begin()); ph.H.erase(ph.H.begin()); ph.K.erase(ph.K.begin());
    ph.rcost.erase(ph.rcost.begin());
    ph.horizon--;
    if (ph.n_path > 0) {
        ph.g.erase(ph.g.begin(), ph.g.begin() + ph.n_path);
        ph.reb.erase(ph.reb.begin(), ph.reb.begin() + ph.n_path);
    }
}

// SinglePhase::push_back_default: Trajectory::push_back_state(X.back()) (TrajectoryManagement.cpp:178-207) +
// PathConstraintBase::push_back (ConstraintsBase.h:276-280: the new stage copies the LAST stage's ReB parameters)
void Problem::phase_push_back_default(Phase& ph) {
    const Vec24 xb = ph.X.back();
    ph.Xbar.push_back(xb); ph.X.push_back(xb);
    ph.Ubar.push_back(Vec24{}); ph.U.push_back(Vec24{});
    ph.Xsim.push_back(Vec24{}); ph.Defect.push_back(Vec24{}); ph.Defect_bar.push_back(Vec24{});
    ph.A.push_back(Mat24{}); ph.B.push_back(Mat24{});
    ph.dU.push_back(Vec24{}); ph.G.push_back(Vec24{}); ph.H.push_back(Mat24{}); ph.K.push_back(Mat24{}); ph.dX.push_back(Vec24{});
    RCost rc; rc.zero();
    ph.rcost.push_back(rc);
    ph.horizon++;
    if (ph.n_path > 0) {
        for (int i = 0; i < ph.n_path; ++i) ph.g.push_back(0.0);
        const size_t last = ph.reb.size() - ph.n_path;
        for (int i = 0; i < ph.n_path; ++i) ph.reb.push_back(ph.reb[last + i]);
    }
}

// HKDProblem::update (HKDProblem.cpp:117-222), nsteps_between_mpc = 1
void Problem::update() {
    // update the reference by one simulation time step
    ref.step(dt_sim);
    const float new_start_time = ref.start_time;
    const float new_end_time = ref.end_time;
    // ---- front end ----
    phases.front().start_time += dt_sim;
    if (approx_leq_scalar(phases.front().end_time, new_start_time)) {
        phases.erase(phases.begin());  // pop_front_phase: the first phase has shrunk to a point
    } else {
        phase_pop_front(phases.front());
        phases.front().start_time = new_start_time;
    }
    // ---- back end ----
    const int* cnew = ref.contact_at_t(new_end_time - new_start_time);
    int new_contact[4];
    std::copy(cnew, cnew + 4, new_contact);
    bool contact_change = false;
    for (int l = 0; l < 4; ++l) contact_change = contact_change || (new_contact[l] != phases.back().contact[l]);
    if (contact_change && phases.back().reach_end) {
        // grow the multi-phase problem by a new phase (zero-initialised trajectory, empty shooting set, no reset map yet)
        Phase ph;
        ph.start_time = phases.back().end_time;
        ph.end_time = new_end_time;
        const int hz = (int)std::round((ph.end_time - ph.start_time) / dt_sim);
        ph.reach_end = false;
        std::copy(new_contact, new_contact + 4, ph.contact);
        phases.push_back(ph);
        create_phase(phases.back(), hz);
    } else {
        // grow the last phase by one time step
        phases.back().end_time = new_end_time;
        if (contact_change) phases.back().reach_end = true;
        phase_push_back_default(phases.back());
    }
    if (phases.back().reach_end) add_tconstr((int)phases.size() - 1);
    // ---- shooting configuration ----
    const int n = (int)phases.size();
    for (int i = 0; i < n; ++i) {
        phases[i].t_offset = phases[i].start_time - phases[0].start_time;
        // reset_params(): a no-op in the reference (ConstraintsBase.h:165-167,341-348): ReB / AL parameters persist
        if ((i == n - 1 && phases[i].horizon > 2) || i < n - 1) phases[i].ss_size = phases[i].horizon + 1;
        phases.front().Ubar[0].zero();
    }
}

// HKDMPC.cpp:44-54
void Problem::default_x0(Vec24& x) const {
    double body[12] = {0, 0, 0, 0, 0, 0.2486, 0, 0, 0, 0, 0, 0};
    double qJ[12] = {0, -0.8, 1.6, 0, -0.8, 1.6, 0, -0.8, 1.6, 0, -0.8, 1.6};
    double qd[12];
    model.hkd_state(body, body + 3, qJ, phases.front().contact, qd);
    for (int i = 0; i < 12; ++i) { x[i] = body[i]; x[12 + i] = qd[i]; }
}

// HKDReference.cpp:8-57 (Q12): foot placement if the SAMPLE's contact flag is set, else joint angle.
void Problem::reference_at_t(float t, double xr[24], double ur[24], int* sample_idx) const {
    int k = ref.index_at_t(t);
    if (sample_idx) *sample_idx = k;
    const double* bs = ref.body_state(k);
    for (int i = 0; i < 12; ++i) xr[i] = bs[i];
    const int* c = ref.contact(k);
    for (int l = 0; l < 4; ++l)
        for (int j = 0; j < 3; ++j)
            xr[12 + 3 * l + j] = (c[l] > 0) ? ref.foot(k)[3 * l + j] : ref.qJ(k)[3 * l + j];
    const double* f = ref.grf(k);
    for (int i = 0; i < 12; ++i) { ur[i] = f[i]; ur[12 + i] = 0.0; }  // qJd is never loaded
}

// ---------------------------------------------------------------------------
// Reset map (HKDReset.h:41-136)
// ---------------------------------------------------------------------------
void Problem::resetmap(const Phase& ph, const Vec24& x, Vec24& xnext) const {
    const double* eul = &x.v[0];
    const double* pos = &x.v[3];
    double qd[12];
    for (int i = 0; i < 12; ++i) qd[i] = x[12 + i];
    for (int l = 0; l < 4; ++l) {
        if (ph.contact[l] && !ph.next_contact[l]) { qd[3 * l] = 0.0; qd[3 * l + 1] = -0.8; qd[3 * l + 2] = 1.7; }
        if (!ph.contact[l] && ph.next_contact[l]) {
            double qleg[3] = {qd[3 * l], qd[3 * l + 1], qd[3 * l + 2]};
            double pf[3];
            model.foot_position(pos, eul, qleg, l, pf);
            qd[3 * l] = 1.0 * pf[0]; qd[3 * l + 1] = 1.0 * pf[1]; qd[3 * l + 2] = 0.0 * pf[2];
        }
    }
    for (int i = 0; i < 12; ++i) { xnext[i] = x[i]; xnext[12 + i] = qd[i]; }
}

void Problem::resetmap_partial(const Phase& ph, const Vec24& x, Mat24& Px) const {
    Px.identity();
    const double* eul = &x.v[0];
    const double* pos = &x.v[3];
    for (int l = 0; l < 4; ++l) {
        if (ph.contact[l] && !ph.next_contact[l])
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 24; ++c) Px(12 + 3 * l + r, c) = 0.0;
        if (!ph.contact[l] && ph.next_contact[l]) {
            double qleg[3] = {x[12 + 3 * l], x[13 + 3 * l], x[14 + 3 * l]};
            double J[54];
            model.foot_jacobian(pos, eul, qleg, l, J);
            const double cmap[3] = {1.0, 1.0, 0.0};
            for (int c = 0; c < 18; ++c)
                for (int r = 0; r < 3; ++r) J[r + 3 * c] = cmap[r] * J[r + 3 * c];
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 3; ++c) {
                    Px(12 + 3 * l + r, c) = J[r + 3 * (3 + c)];      // d/d eul
                    Px(12 + 3 * l + r, 3 + c) = J[r + 3 * c];        // d/d pos
                }
                for (int c = 0; c < 12; ++c) Px(12 + 3 * l + r, 12 + c) = J[r + 3 * (6 + c)];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Constraints (HKDConstraints.cpp:7-171, ConstraintsBase.h:191-202,366-373)
// ---------------------------------------------------------------------------
void Problem::grf_violation(Phase& ph, int k) {
    if (ph.n_path == 0) return;
    const double mu = cparams.mu;
    const Vec24& u = ph.U[k];
    double* g = &ph.g[(size_t)k * ph.n_path];
    for (int s = 0; s < ph.n_stance; ++s) {
        const int l = ph.stance_legs[s];
        const double fx = u[3 * l], fy = u[3 * l + 1], fz = u[3 * l + 2];
        g[5 * s + 0] = fz;
        g[5 * s + 1] = -fx + mu * fz;
        g[5 * s + 2] = fx + mu * fz;
        g[5 * s + 3] = -fy + mu * fz;
        g[5 * s + 4] = fy + mu * fz;
    }
    // update_max_violation(k)
    if (k == 0) ph.path_max_violation = 0;
    double mk = 0;
    for (int i = 0; i < ph.n_path; ++i) mk = std::min(mk, g[i]);
    ph.path_max_violation = std::min(ph.path_max_violation, mk);
}

void Problem::td_violation(Phase& ph) {
    const Vec24& x = ph.X[ph.horizon];
    for (auto& td : ph.tds) {  // ConstraintContainer::compute_terminal_constraints: every constraint object
        for (int i = 0; i < td.n_td; ++i) {
            const int l = td.td_legs[i];
            double qleg[3] = {x[12 + 3 * l], x[13 + 3 * l], x[14 + 3 * l]};
            double pf[3];
            model.foot_position(&x.v[3], &x.v[0], qleg, l, pf);
            td.h[i] = pf[2] - 0.0;
        }
        td.max_violation = 0.0;
        for (int i = 0; i < td.n_td; ++i) td.max_violation = std::max(td.max_violation, std::fabs(td.h[i]));
    }
}

void Problem::td_partial(Phase& ph) {
    const Vec24& x = ph.X[ph.horizon];
    for (auto& td : ph.tds)
        for (int i = 0; i < td.n_td; ++i) {
            const int l = td.td_legs[i];
            double qleg[3] = {x[12 + 3 * l], x[13 + 3 * l], x[14 + 3 * l]};
            double J[54];
            model.foot_jacobian(&x.v[3], &x.v[0], qleg, l, J);
            // Jz = bottom row; hx = [Jz(eul) Jz(pos) 0(6) Jz(qJ)]; entries 6..11 stay zero
            for (int c = 0; c < 3; ++c) { td.hx[i][c] = J[2 + 3 * (3 + c)]; td.hx[i][3 + c] = J[2 + 3 * c]; }
            for (int c = 0; c < 12; ++c) td.hx[i][12 + c] = J[2 + 3 * (6 + c)];
        }
}

// ---------------------------------------------------------------------------
// Costs (HKDCost.h:11-73, HKDCost.cpp:5-66, SinglePhaseInterface.cpp:55-165)
// ---------------------------------------------------------------------------
static inline void tracking_weights(const int contact[4], double Q[24], double Qf[24], double R[24]) {
    const double q[12] = {1, 4, 5, 1, 1, 30, .2, .2, .2, 4, 1, .5};
    for (int i = 0; i < 12; ++i) Q[i] = q[i];
    for (int l = 0; l < 4; ++l)
        for (int j = 0; j < 3; ++j) Q[12 + 3 * l + j] = .2 * (1 - contact[l]);
    const double scale[12] = {1, 1, 2, 1, 1, 20, .3, .3, .3, 1, 3, 1};
    for (int i = 0; i < 24; ++i) Qf[i] = (20 * (i < 12 ? scale[i] : .01)) * Q[i];
    for (int i = 0; i < 12; ++i) { R[i] = .2; R[12 + i] = .1; }
}
static inline void foot_weights(const int contact[4], double Qfoot[12]) {
    for (int l = 0; l < 4; ++l) { Qfoot[3 * l] = 3 * contact[l]; Qfoot[3 * l + 1] = contact[l]; Qfoot[3 * l + 2] = 0; }
    for (int i = 0; i < 12; ++i) Qfoot[i] *= 20;
}

// d_prel of HKDFootPlaceReg (HKDCost.cpp:12-18)
static inline void foot_rel_error(const Vec24& x, const double* body_r, const double* foot_r, double d[12]) {
    for (int l = 0; l < 4; ++l)
        for (int j = 0; j < 3; ++j) {
            double prel = x[12 + 3 * l + j] - x[3 + j];
            double prel_r = foot_r[3 * l + j] - body_r[3 + j];
            d[3 * l + j] = prel - prel_r;
        }
}

void Problem::running_cost(const Phase& ph, int k, RCost& rc) const {
    rc.zero();
    const float t = (float)((double)ph.t_offset + k * ph.dt);
    double xr[24], ur[24];
    int idx;
    reference_at_t(t, xr, ur, &idx);
    double Q[24], Qf[24], R[24], Qfoot[12];
    tracking_weights(ph.contact, Q, Qf, R);
    foot_weights(ph.contact, Qfoot);
    const Vec24& x = ph.X[k];
    const Vec24& u = ph.U[k];
    // tracking
    double l = 0;
    { double s = 0; for (int j = 0; j < 24; ++j) { double dx = x[j] - xr[j]; s += (0.5 * dx * Q[j]) * dx; } l = s; }
    { double s = 0; for (int j = 0; j < 24; ++j) { double du = u[j] - ur[j]; s += (0.5 * du * R[j]) * du; } l += s; }
    l *= ph.dt;
    rc.l += l;
    // foot placement regulariser
    double d[12];
    foot_rel_error(x, ref.body_state(idx), ref.foot(idx), d);
    double lf = 0;
    for (int j = 0; j < 12; ++j) lf += (.5 * d[j] * Qfoot[j]) * d[j];
    lf *= ph.dt;
    rc.l += lf;
}

void Problem::running_cost_par(const Phase& ph, int k, RCost& rc) const {
    const float t = (float)((double)ph.t_offset + k * ph.dt);
    double xr[24], ur[24];
    int idx;
    reference_at_t(t, xr, ur, &idx);
    double Q[24], Qf[24], R[24], Qfoot[12];
    tracking_weights(ph.contact, Q, Qf, R);
    foot_weights(ph.contact, Qfoot);
    const Vec24& x = ph.X[k];
    const Vec24& u = ph.U[k];
    const double dt = ph.dt;
    for (int j = 0; j < 24; ++j) {
        rc.lx[j] += (dt * Q[j]) * (x[j] - xr[j]);
        rc.lu[j] += (dt * R[j]) * (u[j] - ur[j]);
        rc.lxx(j, j) += dt * Q[j];
        rc.luu(j, j) += dt * R[j];
    }
    double d[12];
    foot_rel_error(x, ref.body_state(idx), ref.foot(idx), d);
    for (int l = 0; l < 4; ++l) {
        const double c = ph.contact[l];
        for (int j = 0; j < 3; ++j) {
            const double w = dt * c * Qfoot[3 * l + j];  // one factor c from each dprel_dx
            rc.lx[3 + j] += -(w * d[3 * l + j]);
            rc.lx[12 + 3 * l + j] += w * d[3 * l + j];
            const double wc = w * c;
            rc.lxx(3 + j, 3 + j) += wc;
            rc.lxx(3 + j, 12 + 3 * l + j) += -wc;
            rc.lxx(12 + 3 * l + j, 3 + j) += -wc;
            rc.lxx(12 + 3 * l + j, 12 + 3 * l + j) += wc;
        }
    }
}

void Problem::terminal_cost(const Phase& ph, TCost& tc) const {
    tc.zero();
    const int k = ph.horizon;
    const float t = (float)((double)ph.t_offset + k * ph.dt);
    double xr[24], ur[24];
    int idx;
    reference_at_t(t, xr, ur, &idx);
    double Q[24], Qf[24], R[24], Qfoot[12];
    tracking_weights(ph.contact, Q, Qf, R);
    foot_weights(ph.contact, Qfoot);
    const Vec24& x = ph.X[k];
    double s = 0;
    for (int j = 0; j < 24; ++j) { double dx = x[j] - xr[j]; s += (dx * Qf[j]) * dx; }
    tc.Phi += 0.5 * s;
    double d[12];
    foot_rel_error(x, ref.body_state(idx), ref.foot(idx), d);
    double sf = 0;
    for (int j = 0; j < 12; ++j) sf += (10 * d[j] * Qfoot[j]) * d[j];
    tc.Phi += sf;
}

void Problem::terminal_cost_par(const Phase& ph, TCost& tc) const {
    const int k = ph.horizon;
    const float t = (float)((double)ph.t_offset + k * ph.dt);
    double xr[24], ur[24];
    int idx;
    reference_at_t(t, xr, ur, &idx);
    double Q[24], Qf[24], R[24], Qfoot[12];
    tracking_weights(ph.contact, Q, Qf, R);
    foot_weights(ph.contact, Qfoot);
    const Vec24& x = ph.X[k];
    for (int j = 0; j < 24; ++j) { tc.Phix[j] += Qf[j] * (x[j] - xr[j]); tc.Phixx(j, j) += Qf[j]; }
    double d[12];
    foot_rel_error(x, ref.body_state(idx), ref.foot(idx), d);
    for (int l = 0; l < 4; ++l) {
        const double c = ph.contact[l];
        for (int j = 0; j < 3; ++j) {
            const double w = 20 * c * Qfoot[3 * l + j];
            tc.Phix[3 + j] += -(w * d[3 * l + j]);
            tc.Phix[12 + 3 * l + j] += w * d[3 * l + j];
            const double wc = w * c;
            tc.Phixx(3 + j, 3 + j) += wc;
            tc.Phixx(3 + j, 12 + 3 * l + j) += -wc;
            tc.Phixx(12 + 3 * l + j, 3 + j) += -wc;
            tc.Phixx(12 + 3 * l + j, 12 + 3 * l + j) += wc;
        }
    }
}

// ---------------------------------------------------------------------------
// SinglePhase (SinglePhase.cpp:145-426)
// ---------------------------------------------------------------------------
bool Problem::phase_hybrid_rollout(Phase& ph, double eps, const Options& opt) {
    const int N = ph.horizon;
    ph.Xsim[0] = ph.x_init;
    // the first state is a shooting state iff SS_set is non-empty and starts at 0, whatever option.MS says (SinglePhase.cpp:187-193,
    // Q9); SS_set = {0 .. ss_size-1}: all nodes after initialisation, EMPTY for a phase created by HKDProblem::update until its
    // horizon exceeds 2, one short for a last phase that grew while its horizon was <= 2 (HKDProblem.cpp:212-218)
    if (ph.ss_size > 0) { for (int j = 0; j < 24; ++j) ph.X[0][j] = ph.Xbar[0][j] + eps * ph.dX[0][j]; }
    else ph.X[0] = ph.x_init;
    int k = 0;
    for (k = 0; k < N; ++k) {
        // U = Ubar + eps dU + K (X - Xbar)
        Vec24 dx, Kdx;
        for (int j = 0; j < 24; ++j) dx[j] = ph.X[k][j] - ph.Xbar[k][j];
        matvec(ph.K[k], dx, Kdx);
        for (int j = 0; j < 24; ++j) ph.U[k][j] = (ph.Ubar[k][j] + eps * ph.dU[k][j]) + Kdx[j];
        model.dynamics(ph.X[k].v, ph.U[k].v, ph.dt, ph.contact, ph.Xsim[k + 1].v);
        double nrm2 = 0;
        for (int j = 0; j < 24; ++j) nrm2 += ph.Xsim[k + 1][j] * ph.Xsim[k + 1][j];
        if (std::sqrt(nrm2) > 1e6) return false;  // Q16 (NaN passes)
        if (opt.MS && k + 1 < ph.ss_size) {  // k+1 is a shooting state (SinglePhase.cpp:211-220)
            for (int j = 0; j < 24; ++j) ph.X[k + 1][j] = ph.Xbar[k + 1][j] + eps * ph.dX[k + 1][j];
        } else {
            ph.X[k + 1] = ph.Xsim[k + 1];
        }
        grf_violation(ph, k);
    }
    td_violation(ph);
    for (int i = 0; i <= N; ++i)
        for (int j = 0; j < 24; ++j) ph.Defect[i][j] = ph.Xsim[i][j] - ph.X[i][j];
    return true;
}

void Problem::phase_linear_rollout(Phase& ph, double eps) {
    const int N = ph.horizon;
    ph.dV_1 = 0; ph.dV_2 = 0;
    for (int j = 0; j < 24; ++j) ph.dX[0][j] = ph.dx_init[j] + eps * ph.Defect[0][j];
    for (int k = 0; k < N; ++k) {
        const RCost& rc = ph.rcost[k];
        const Vec24& dxk = ph.dX[k];
        Vec24 duk, Kdx, Adx, Bdu;
        matvec(ph.K[k], dxk, Kdx);
        for (int j = 0; j < 24; ++j) duk[j] = eps * ph.dU[k][j] + Kdx[j];
        matvec(ph.A[k], dxk, Adx);
        matvec(ph.B[k], duk, Bdu);
        for (int j = 0; j < 24; ++j) ph.dX[k + 1][j] = (Adx[j] + Bdu[j]) + eps * ph.Defect[k + 1][j];
        ph.dV_1 += dot(rc.lx, dxk) + dot(rc.lu, duk);
        Vec24 t1;
        matvec_t(rc.lxx, dxk, t1);  // (dx^T Q) then . dx
        ph.dV_2 += dot(t1, dxk);
        matvec_t(rc.luu, duk, t1);
        ph.dV_2 += dot(t1, duk);
        matvec_t(rc.lux, duk, t1);  // (du^T P) . dx  — counted once (Q13)
        ph.dV_2 += dot(t1, dxk);
    }
    const Vec24& dxN = ph.dX[N];
    ph.dV_1 += dot(ph.tcost.Phix, dxN);
    Vec24 t1;
    matvec_t(ph.tcost.Phixx, dxN, t1);
    ph.dV_2 += dot(t1, dxN);
}

void Problem::phase_compute_cost(Phase& ph, const Options& opt) {
    const int N = ph.horizon;
    ph.actual_cost = 0;
    for (int k = 0; k < N; ++k) {
        running_cost(ph, k, ph.rcost[k]);  // zeroes values AND derivatives (Q1)
        if (opt.ReB_active && ph.n_path > 0) {
            // compute_ReB_cost, ConstraintsBase.h:204-222
            double reb_cost = 0;
            const double* g = &ph.g[(size_t)k * ph.n_path];
            const RebParam* p = &ph.reb[(size_t)k * ph.n_path];
            for (int i = 0; i < ph.n_path; ++i) {
                double barr;
                if (g[i] > p[i].delta) barr = -std::log(g[i]);
                else {
                    double z = (g[i] - 2 * p[i].delta) / p[i].delta;
                    barr = .5 * (z * z - 1);
                    barr -= std::log(p[i].delta);
                }
                reb_cost += p[i].eps * barr;
            }
            ph.rcost[k].l += ph.dt * reb_cost;
        }
        ph.actual_cost += ph.rcost[k].l;
    }
    terminal_cost(ph, ph.tcost);
    if (opt.AL_active) {
        for (auto& td : ph.tds) {  // update_terminal_cost_with_tconstr, SinglePhase.cpp:402-411: one AL_cost per constraint object
            double al_cost = 0;    // compute_AL_cost, ConstraintsBase.h:374-385
            for (int i = 0; i < td.n_td; ++i) {
                al_cost += 0.5 * td.al[i].sigma * td.h[i] * td.h[i];
                al_cost += td.al[i].lambda * td.h[i];
            }
            ph.tcost.Phi += al_cost;
        }
    }
    ph.actual_cost += ph.tcost.Phi;
}

void Problem::phase_LQ_approximation(Phase& ph, const Options& opt) {
    const int N = ph.horizon;
    const double mu = cparams.mu;
    for (int k = 0; k < N; ++k) {
        model.dynamics_partial(ph.X[k].v, ph.U[k].v, ph.dt, ph.contact, ph.A[k].m, ph.B[k].m);
        running_cost_par(ph, k, ph.rcost[k]);  // adds onto the zeroed derivatives
        if (opt.ReB_active && ph.n_path > 0) {
            // compute_ReB_partials, ConstraintsBase.h:224-263 (only gu is non-zero for the GRF rows)
            Vec24 grad_u; grad_u.zero();
            Mat24 hess_u; hess_u.zero();
            const double* g = &ph.g[(size_t)k * ph.n_path];
            const RebParam* p = &ph.reb[(size_t)k * ph.n_path];
            for (int s = 0; s < ph.n_stance; ++s) {
                const int l = ph.stance_legs[s];
                const double rows[5][3] = {{0, 0, 1}, {-1, 0, mu}, {1, 0, mu}, {0, -1, mu}, {0, 1, mu}};
                for (int r = 0; r < 5; ++r) {
                    const int i = 5 * s + r;
                    double bd, bdd;
                    if (g[i] > p[i].delta) { bd = -1.0 / g[i]; bdd = std::pow(g[i], -2); }
                    else { bd = (g[i] - 2 * p[i].delta) / p[i].delta / p[i].delta; bdd = std::pow(p[i].delta, -2); }
                    for (int a = 0; a < 3; ++a) grad_u[3 * l + a] += p[i].eps * bd * rows[r][a];
                    for (int b = 0; b < 3; ++b)
                        for (int a = 0; a < 3; ++a)
                            hess_u(3 * l + a, 3 * l + b) += p[i].eps * (bdd * rows[r][a] * rows[r][b] + bd * 0.0);
                }
            }
            for (int j = 0; j < 24; ++j) ph.rcost[k].lu[j] += ph.dt * grad_u[j];
            for (int j = 0; j < 576; ++j) ph.rcost[k].luu.m[j] += ph.dt * hess_u.m[j];
        }
    }
    terminal_cost_par(ph, ph.tcost);
    if (opt.AL_active && !ph.tds.empty()) {
        td_partial(ph);
        // compute_AL_partials, ConstraintsBase.h:386-399 (Q3: Hessian weight sigma(1+h)+lambda), one gradient / Hessian per
        // constraint object, added to the terminal cost in turn (SinglePhase.cpp:414-426)
        for (auto& td : ph.tds) {
            Vec24 grad; grad.zero();
            Mat24 hess; hess.zero();
            for (int i = 0; i < td.n_td; ++i) {
                const double wg = td.al[i].sigma * td.h[i] + td.al[i].lambda;
                const double wh = td.al[i].sigma * (1 + td.h[i]) + td.al[i].lambda;
                for (int a = 0; a < 24; ++a) grad[a] += wg * td.hx[i][a];
                for (int b = 0; b < 24; ++b)
                    for (int a = 0; a < 24; ++a) hess(a, b) += wh * (td.hx[i][a] * td.hx[i][b]);
            }
            for (int a = 0; a < 24; ++a) ph.tcost.Phix[a] += grad[a];
            for (int j = 0; j < 576; ++j) ph.tcost.Phixx.m[j] += hess.m[j];
        }
    }
}

bool Problem::phase_backward_sweep(Phase& ph, double reg, const Vec24& Gprime, const Mat24& Hprime) {
    const int N = ph.horizon;
    bool success = true;
    for (int j = 0; j < 24; ++j) ph.G[N][j] = ph.tcost.Phix[j] + Gprime[j];
    for (int j = 0; j < 576; ++j) ph.H[N].m[j] = ph.tcost.Phixx.m[j] + Hprime.m[j];
    ph.dV_1 = 0; ph.dV_2 = 0;
    Mat24 AtH, BtH, Qxx, Quu, Qux, Quu_s, inv1, Quu_inv, tmp, QuxT_Qi;
    Vec24 Qx, Qu, Gnext, t;
    for (int k = N - 1; k >= 0; --k) {
        const RCost& rc = ph.rcost[k];
        const Mat24& Ak = ph.A[k];
        const Mat24& Bk = ph.B[k];
        const Mat24& Hn = ph.H[k + 1];
        // Gnext = G[k+1] + H[k+1] * Defect[k+1]     (Q10)
        matvec(Hn, ph.Defect[k + 1], t);
        for (int j = 0; j < 24; ++j) Gnext[j] = ph.G[k + 1][j] + t[j];
        matvec_t(Ak, Gnext, t); for (int j = 0; j < 24; ++j) Qx[j] = rc.lx[j] + t[j];
        matvec_t(Bk, Gnext, t); for (int j = 0; j < 24; ++j) Qu[j] = rc.lu[j] + t[j];
        // (A^T H) A etc., left-to-right as Eigen associates
        matmul_tn(Ak, Hn, AtH);
        matmul_tn(Bk, Hn, BtH);
        matmul(AtH, Ak, tmp); for (int j = 0; j < 576; ++j) Qxx.m[j] = rc.lxx.m[j] + tmp.m[j];
        matmul(BtH, Bk, tmp); for (int j = 0; j < 576; ++j) Quu.m[j] = rc.luu.m[j] + tmp.m[j];
        matmul(BtH, Ak, tmp); for (int j = 0; j < 576; ++j) Qux.m[j] = rc.lux.m[j] + tmp.m[j];
        for (int j = 0; j < 24; ++j) { Qxx(j, j) += 1.0 * reg; Quu(j, j) += 1.0 * reg; }
        // PD test on Quu - 1e-9 I (Q7)
        Quu_s = Quu;
        for (int j = 0; j < 24; ++j) Quu_s(j, j) -= 1.0 * 1e-9;
        if (!ldlt_is_positive(Quu_s)) { success = false; break; }
        inverse_partial_piv_lu(Quu, inv1);
        for (int j = 0; j < 24; ++j)
            for (int i = 0; i < 24; ++i) Quu_inv(i, j) = (inv1(i, j) + inv1(j, i)) / 2;
        for (int j = 0; j < 24; ++j)
            for (int i = 0; i < 24; ++i) tmp(i, j) = (Qxx(i, j) + Qxx(j, i)) / 2;
        Qxx = tmp;
        // dU = -Quu_inv Qu ; K = -Quu_inv Qux
        matvec(Quu_inv, Qu, t); for (int j = 0; j < 24; ++j) ph.dU[k][j] = -t[j];
        matmul(Quu_inv, Qux, tmp); for (int j = 0; j < 576; ++j) ph.K[k].m[j] = -tmp.m[j];
        // G = Qx - (Qux^T Quu_inv) Qu ; H = Qxx - (Qux^T Quu_inv) Qux
        matmul_tn(Qux, Quu_inv, QuxT_Qi);
        matvec(QuxT_Qi, Qu, t); for (int j = 0; j < 24; ++j) ph.G[k][j] = Qx[j] - t[j];
        matmul(QuxT_Qi, Qux, tmp); for (int j = 0; j < 576; ++j) ph.H[k].m[j] = Qxx.m[j] - tmp.m[j];
        double dV_k = -dot(Qu, ph.dU[k]);
        ph.dV_1 -= dV_k;
        ph.dV_2 += dV_k;
    }
    // runs even after a failed stage (SinglePhase.cpp:365)
    matvec(ph.H[0], ph.Defect[0], t);
    for (int j = 0; j < 24; ++j) ph.G[0][j] += t[j];
    return success;
}

// ---------------------------------------------------------------------------
// MultiPhaseDDP (MultiPhaseDDP.cpp)
// ---------------------------------------------------------------------------
void Problem::linear_rollout(double eps, const Options& opt) {
    (void)opt;
    Vec24 dx_init; dx_init.zero();
    dV_1 = 0; dV_2 = 0;
    const int n = (int)phases.size();
    for (int i = 0; i < n; ++i) {
        if (i > 0) {
            Mat24 Px;
            resetmap_partial(phases[i - 1], phases[i - 1].X.back(), Px);
            matvec(Px, phases[i - 1].dX.back(), dx_init);
        }
        phases[i].dx_init = dx_init;
        phase_linear_rollout(phases[i], eps);
        dV_1 += phases[i].dV_1;
        dV_2 += phases[i].dV_2;
    }
}

bool Problem::hybrid_rollout(double eps, const Options& opt) {
    actual_cost = 0; max_pconstr = 0; max_tconstr = 0;
    Vec24 xinit = x0;
    bool success = true;
    const int n = (int)phases.size();
    for (int i = 0; i < n; ++i) {
        if (i > 0) resetmap(phases[i - 1], phases[i - 1].X.back(), xinit);
        phases[i].x_init = xinit;
        if (!phase_hybrid_rollout(phases[i], eps, opt)) { success = false; break; }
        // ConstraintContainer::get_max_{p,t}constrs: 0 when the phase has no such constraint
        double mp = phases[i].n_path > 0 ? std::min(0.0, phases[i].path_max_violation) : 0.0;
        double mt = 0.0;
        for (auto& td : phases[i].tds) mt = std::max(mt, td.max_violation);
        max_pconstr = std::min(max_pconstr, mp);
        max_tconstr = std::max(max_tconstr, mt);
    }
    return success;
}

void Problem::compute_cost(const Options& opt) {
    actual_cost = 0;
    for (auto& ph : phases) { phase_compute_cost(ph, opt); actual_cost += ph.actual_cost; }
}

void Problem::LQ_approximation(const Options& opt) {
    for (auto& ph : phases) phase_LQ_approximation(ph, opt);
}

double Problem::measure_dynamics_feasibility() {
    double f = 0;
    for (auto& ph : phases) {
        double s = 0;
        for (auto& d : ph.Defect) { double q = 0; for (int j = 0; j < 24; ++j) q += d[j] * d[j]; s += q; }
        f += s;
    }
    return std::sqrt(f);
}

bool Problem::line_search(const Options& opt, int* n_trials, double* eps_out) {
    double eps = 1;
    const double merit_prev = merit;
    const double feas_prev = feas;
    bool success = false;
    int trials = 0;
    while (eps > 1e-3) {  // 1, .1, .010000000000000002, .0010000000000000002 (Q6)
        bool rollout_success = hybrid_rollout(eps, opt);
        compute_cost(opt);
        feas = measure_dynamics_feasibility();
        merit = actual_cost + merit_rho * feas;
        ++trials;
        double exp_cost_change = eps * dV_1 + 0.5 * eps * eps * dV_2;
        double exp_merit_change = exp_cost_change - eps * merit_rho * feas_prev;
        if ((merit <= merit_prev + opt.gamma * exp_merit_change) && rollout_success) { success = true; break; }
        eps *= opt.alpha;
    }
    if (n_trials) *n_trials = trials;
    if (eps_out) *eps_out = success ? eps : 0.0;
    return success;
}

bool Problem::backward_sweep(double regularization) {
    const int n = (int)phases.size();
    dV_1 = 0; dV_2 = 0;
    for (int i = n - 1; i >= 0; --i) {
        Vec24 Gp; Gp.zero();
        Mat24 Hp; Hp.zero();
        if (i <= n - 2) {
            Mat24 Px, PtH;
            resetmap_partial(phases[i], phases[i].X.back(), Px);
            // impact_aware_step: G = Px^T G ; H = (Px^T H) Px
            matvec_t(Px, phases[i + 1].G[0], Gp);
            matmul_tn(Px, phases[i + 1].H[0], PtH);
            matmul(PtH, Px, Hp);
        }
        if (!phase_backward_sweep(phases[i], regularization, Gp, Hp)) return false;
        dV_1 += phases[i].dV_1;
        dV_2 += phases[i].dV_2;
    }
    return true;
}

bool Problem::backward_sweep_regularized(double& regularization, const Options& opt, int* n_sweeps) {
    bool success = false;
    int iter = 0;
    while (!success) {
        ++iter;
        success = backward_sweep(regularization);
        if (success) break;
        regularization = std::max(regularization * opt.update_regularization, 1e-03);
        if (regularization > 1e2) break;
    }
    if (n_sweeps) *n_sweeps = iter;
    regularization = regularization / 20;
    if (regularization < 1e-06) regularization = 0;
    return success;
}

void Problem::update_nominal_trajectory() {
    for (auto& ph : phases) { ph.Xbar = ph.X; ph.Ubar = ph.U; ph.Defect_bar = ph.Defect; }
}

void Problem::update_AL_params(const Options& opt) {
    for (auto& ph : phases)
        for (auto& td : ph.tds)
            for (int i = 0; i < td.n_td; ++i) {  // TerminalConstraintBase::update_params, ConstraintsBase.h:349-365
                if (std::fabs(td.h[i]) < opt.tconstr_thresh) continue;
                if (std::fabs(td.h[i]) > 0.005) {
                    td.al[i].sigma *= opt.update_penalty;
                    td.al[i].sigma = std::min(td.al[i].sigma, td.al[i].sigma_max);
                } else {
                    td.al[i].lambda += td.h[i] * td.al[i].sigma;
                }
            }
}

void Problem::update_REB_params(const Options& opt) {
    for (auto& ph : phases)
        for (size_t i = 0; i < ph.g.size(); ++i) {  // PathConstraintBase::update_params, ConstraintsBase.h:168-183
            if (ph.g[i] > -opt.pconstr_thresh) continue;
            ph.reb[i].eps *= opt.update_ReB;
            ph.reb[i].delta *= opt.update_relax;
            ph.reb[i].delta = std::fmax(ph.reb[i].delta, ph.reb[i].delta_min);
        }
}

void Problem::solve(const Options& option, SolveResult& out) {
    out = SolveResult();
    int iter = 0, iter_ou = 0, iter_in = 0;
    double cost_prev = 0, merit_prev = 0;
    bool success = true;
    actual_cost = 0; max_pconstr = 0; max_pconstr_prev = 0; max_tconstr = 0; max_tconstr_prev = 0;

    hybrid_rollout(0, option);
    update_nominal_trajectory();
    compute_cost(option);
    feas = measure_dynamics_feasibility();
    out.cost0 = actual_cost; out.feas0 = feas;
    out.cost_buffer.push_back((float)actual_cost);
    out.dyn_feas_buffer.push_back((float)feas);
    out.eqn_feas_buffer.push_back((float)max_tconstr);
    out.ineq_feas_buffer.push_back((float)max_pconstr);

    int status = 2;
    while (iter_ou < option.max_AL_iter) {
        iter_ou++;
        max_tconstr_prev = max_tconstr;
        max_pconstr_prev = max_pconstr;
        double regularization = 0;
        iter_in = 0;
        while (iter_in < option.max_DDP_iter) {
            compute_cost(option);
            feas = measure_dynamics_feasibility();
            iter_in++; iter++;
            IterRecord rec{};
            rec.outer = iter_ou; rec.inner = iter_in;
            rec.cost_before = actual_cost; rec.feas_before = feas;

            LQ_approximation(option);
            int nsw = 0;
            success = backward_sweep_regularized(regularization, option, &nsw);
            out.n_sweeps += nsw;
            rec.n_sweeps = nsw;
            rec.reg_used = regularization;  // value AFTER the /20 decay (what the next iteration starts from)
            if (!success) { out.trace.push_back(rec); goto bad_solve; }

            if (option.MS) linear_rollout(1.0, option);
            {
                double dV_abs = std::fabs(dV_1 + 0.5 * dV_2);
                merit_rho = (feas > option.dynamics_feas_thresh) ? dV_abs / ((1 - option.merit_scale) * feas) + option.merit_offset : 0;
                merit = actual_cost + merit_rho * feas;
                cost_prev = actual_cost;
                merit_prev = merit;
                rec.dV_1 = dV_1; rec.dV_2 = dV_2; rec.merit_rho = merit_rho;
                if ((dV_abs < option.cost_thresh) && (feas <= option.dynamics_feas_thresh)) {
                    rec.eps_accepted = -1; rec.n_trials = 0;
                    rec.cost_after = actual_cost; rec.feas_after = feas; rec.max_tconstr = max_tconstr; rec.max_pconstr = max_pconstr;
                    out.trace.push_back(rec);
                    break;
                }
            }
            {
                int ntr = 0; double eps_acc = 0;
                if (line_search(option, &ntr, &eps_acc)) update_nominal_trajectory();
                else { actual_cost = cost_prev; merit = merit_prev; }  // Q2: nothing else is rolled back
                rec.eps_accepted = eps_acc; rec.n_trials = ntr;
            }
            rec.cost_after = actual_cost; rec.feas_after = feas; rec.max_tconstr = max_tconstr; rec.max_pconstr = max_pconstr;
            out.trace.push_back(rec);
            if ((std::fabs((cost_prev - actual_cost) / cost_prev) < option.cost_thresh) && (feas <= option.dynamics_feas_thresh)) break;
            out.cost_buffer.push_back((float)actual_cost);
            out.dyn_feas_buffer.push_back((float)feas);
            out.eqn_feas_buffer.push_back((float)max_tconstr);
            out.ineq_feas_buffer.push_back((float)max_pconstr);
        }
        if (option.AL_active) update_AL_params(option);
        if (option.ReB_active) update_REB_params(option);
        if (max_tconstr < option.tconstr_thresh && std::fabs(max_pconstr) < option.pconstr_thresh && feas <= option.dynamics_feas_thresh) { status = 0; break; }
        if (std::fabs(max_tconstr - max_tconstr_prev) < 0.0001 && std::fabs(max_pconstr - max_pconstr_prev) < 0.0001 && feas <= option.dynamics_feas_thresh) { status = 1; break; }
    }
    out.status = status;
bad_solve:
    if (!success) out.status = 3;
    out.n_iter = iter; out.n_outer = iter_ou;
    out.cost = actual_cost; out.feas = feas; out.max_tconstr = max_tconstr; out.max_pconstr = max_pconstr;
    (void)merit_prev;
}

}  // namespace oracle
