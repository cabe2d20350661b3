// Host-side problem assembly for the batched HS-DDP solver (CPU only, no CUDA).
//
// Mirrors, for the inputs of the hot path only:
//   QuadReference::load_top_level_data / initialize / get_a_reference_ptr_at_t / get_contact_at_t
//                                     Reference/QuadReference.cpp:6-26,65-100,129-285
//   HKDSinglePhaseReference::get_reference_at_t        HKDMPC/HKD-TrajOpt/HKDReference.cpp:8-57
//   HKDProblem::initialization / add_tconstr_one_phase HKDMPC/HKD-TrajOpt/HKDProblem.cpp:15-111,268-310
//   compute_hkd_state / default initial condition      HKDModel.h:65-96, HKDMPC/HKDMPC.cpp:44-54
// and flattens the result into the POD `hsddp_schedule` of include/hsddp_b200.h.
// All time arithmetic is done in `float` exactly where the reference uses float
// (SURVEY.md Q11): t_offset, the phase-split loop, the nearest-sample lookup.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include "../../include/hsddp_b200.h"
#include "../csrc/hkd_model.cuh"

struct hkd_gait {
    int n = 0;
    float dt = 0.f;
    std::vector<double> body_state, qJ, foot, grf;  // n x 12, values exactly representable in float
    std::vector<int> contact;                       // n x 4
};

namespace {

inline bool approx_eq(float a, float b) { float tol = 1e-6f; return std::fabs(a - b) <= tol; }  // HSDDP_Utils.h:46-56
inline bool approx_leq(float a, float b) { return a < b || approx_eq(a, b); }
inline bool approx_geq(float a, float b) { return a > b || approx_eq(a, b); }

struct Window {
    const hkd_gait* g;
    int k0;
    int sz;  // round(plan/dt) + 1; the window holds sz + 1 samples
    int index_at(float t) const {  // QuadReference.cpp:65-80
        int k = (int)std::floor(t / g->dt);
        float rem = t - (float)k * g->dt;
        if ((double)rem > 0.5 * (double)g->dt) k++;
        if (k > sz) k = sz;
        return k;
    }
    const int* contact(int k) const { return &g->contact[4 * (size_t)(k0 + k)]; }
};

void parse_floats(const std::string& line, double* dst, int n) {
    std::stringstream ss(line);
    std::string w;
    int i = 0;
    while (ss >> w) {
        dst[i] = (double)std::stof(w);
        if (++i >= n) break;
    }
}

unsigned mask_of(const int32_t c[4]) { return (c[0] ? 1u : 0u) | (c[1] ? 2u : 0u) | (c[2] ? 4u : 0u) | (c[3] ? 8u : 0u); }

}  // namespace

extern "C" {

int hkd_gait_create(int n, float dt, const float* body_state, const float* qJ, const float* foot_placements,
                    const float* grf, const int32_t* contact, hkd_gait** out) {
    if (!out || n <= 0 || !(dt > 0.f) || !body_state || !qJ || !foot_placements || !grf || !contact) return HSDDP_ERR_ARG;
    hkd_gait* g = nullptr;
    try {  // (no C++ exception may cross the C ABI)
        g = new hkd_gait();
        g->n = n;
        g->dt = dt;
        g->body_state.assign(body_state, body_state + 12 * (size_t)n);
        g->qJ.assign(qJ, qJ + 12 * (size_t)n);
        g->foot.assign(foot_placements, foot_placements + 12 * (size_t)n);
        g->grf.assign(grf, grf + 12 * (size_t)n);
        g->contact.assign(contact, contact + 4 * (size_t)n);
    } catch (...) {
        delete g;
        return HSDDP_ERR_ARG;
    }
    *out = g;
    return HSDDP_OK;
}

// Text loader: keys matched by substring in a fixed order; every number through stof.
int hkd_gait_load(const char* path, hkd_gait** out) {
    if (!path || !out) return HSDDP_ERR_ARG;
    hkd_gait* g = nullptr;
    try {  // std::stof / std::stoi throw on malformed text, the vectors on allocation failure: report HSDDP_ERR_IO instead
    std::ifstream f(path);
    if (!f.is_open()) return HSDDP_ERR_IO;
    g = new hkd_gait();
    std::string line;
    double body[12] = {0}, qj[12] = {0}, foot[12] = {0}, grf[12] = {0}, sdur[12] = {0};
    int contact[4] = {0, 0, 0, 0};
    while (std::getline(f, line)) {
        if (line == "dt") { std::getline(f, line); g->dt = std::stof(line); continue; }
        if (line.find("body_state") != std::string::npos) {
            std::memset(body, 0, sizeof body); std::memset(qj, 0, sizeof qj); std::memset(foot, 0, sizeof foot);
            std::memset(grf, 0, sizeof grf); std::memset(contact, 0, sizeof contact);
            std::getline(f, line); parse_floats(line, body, 12); continue;
        }
        if (line.find("qJ") != std::string::npos) { std::getline(f, line); parse_floats(line, qj, 12); continue; }
        if (line.find("foot_placements") != std::string::npos) { std::getline(f, line); parse_floats(line, foot, 12); continue; }
        if (line.find("grf") != std::string::npos) { std::getline(f, line); parse_floats(line, grf, 12); continue; }
        if (line.find("torque") != std::string::npos) { std::getline(f, line); continue; }
        if (line.find("contact") != std::string::npos) {
            std::getline(f, line);
            std::stringstream ss(line); std::string w; int i = 0;
            while (ss >> w) { if (i < 4) contact[i] = std::stoi(w); if (++i >= 12) break; }
            continue;
        }
        if (line.find("status_dur") != std::string::npos) {
            std::getline(f, line); parse_floats(line, sdur, 12);
            g->body_state.insert(g->body_state.end(), body, body + 12);
            g->qJ.insert(g->qJ.end(), qj, qj + 12);
            g->foot.insert(g->foot.end(), foot, foot + 12);
            g->grf.insert(g->grf.end(), grf, grf + 12);
            g->contact.insert(g->contact.end(), contact, contact + 4);
            g->n++;
        }
    }
    if (g->n == 0 || !(g->dt > 0.f)) { delete g; return HSDDP_ERR_IO; }
    } catch (...) {
        delete g;
        return HSDDP_ERR_IO;
    }
    *out = g;
    return HSDDP_OK;
}

int hkd_gait_size(const hkd_gait* g) { return g ? g->n : 0; }
void hkd_gait_destroy(hkd_gait* g) { delete g; }

int hkd_schedule_build(const hkd_gait* g, int window_start, float plan_duration, hsddp_schedule* out) {
    if (!g || !out || window_start < 0 || !(plan_duration > 0.f) || !(g->dt > 0.f)) return HSDDP_ERR_ARG;
    std::memset(out, 0, sizeof *out);
    const float dt_sim = 0.01f;          // HKDMPC.cpp:28
    const float dt_mpc = dt_sim * 1;     // nsteps_between_mpc = 1, HKDProblem.h:104-108
    Window w{g, window_start, (int)std::round(plan_duration / g->dt) + 1};
    if (window_start + w.sz >= g->n) return HSDDP_ERR_ARG;  // QuadReference::initialize copies sz+1 samples

    // ---- phase split, HKDProblem.cpp:26-68 ----
    int contact_prev[4], contact_cur[4];
    float phase_start = 0.f, t = 0.f;
    std::memcpy(contact_prev, w.contact(w.index_at(t)), sizeof contact_prev);
    int n_phases = 0;
    while (approx_leq(t, plan_duration)) {
        std::memcpy(contact_cur, w.contact(w.index_at(t)), sizeof contact_cur);
        bool change = false;
        for (int l = 0; l < 4; ++l) change = change || (contact_cur[l] != contact_prev[l]);
        if (change || approx_geq(t, plan_duration)) {
            if (n_phases >= HSDDP_MAX_PHASES) return HSDDP_ERR_UNSUPPORTED;
            const float phase_end = t;
            out->horizon[n_phases] = (int)std::round((phase_end - phase_start) / dt_sim);
            out->start_time[n_phases] = phase_start;
            for (int l = 0; l < 4; ++l) out->contact[n_phases][l] = contact_prev[l];
            ++n_phases;
            std::memcpy(contact_prev, contact_cur, sizeof contact_prev);
            phase_start = phase_end;
        }
        t += dt_sim;
    }
    out->n_phases = n_phases;
    out->dt = (double)dt_sim;
    int n_stages = 0;
    for (int i = 0; i < n_phases; ++i) {
        if (out->horizon[i] < 1) return HSDDP_ERR_UNSUPPORTED;
        n_stages += out->horizon[i];
    }
    if (n_stages > HSDDP_MAX_STAGES) return HSDDP_ERR_UNSUPPORTED;
    out->n_stages = n_stages;
    out->n_nodes = n_stages + n_phases;
    // ---- contact after each phase (reset map / touchdown wiring), HKDProblem.cpp:268-299, Q17 ----
    for (int i = 0; i < n_phases; ++i) {
        const int* cn = (i < n_phases - 1) ? nullptr : w.contact(w.index_at(plan_duration + dt_mpc));
        for (int l = 0; l < 4; ++l) out->next_contact[i][l] = cn ? cn[l] : out->contact[i + 1][l];
    }
    // ---- per-node reference rows ----
    const size_t nn = (size_t)out->n_nodes;
    out->xr = (double*)std::calloc(nn * 24, sizeof(double));
    out->ur = (double*)std::calloc(nn * 24, sizeof(double));
    out->prel_r = (double*)std::calloc(nn * 12, sizeof(double));
    out->xinit = (double*)std::calloc(nn * 24, sizeof(double));
    if (!out->xr || !out->ur || !out->prel_r || !out->xinit) { hkd_schedule_free(out); return HSDDP_ERR_ARG; }
    auto fill_state_ref = [&](int k, double* x) {  // HKDReference.cpp:33-56 (Q12)
        const double* bs = &g->body_state[12 * (size_t)(w.k0 + k)];
        const int* c = w.contact(k);
        for (int i = 0; i < 12; ++i) x[i] = bs[i];
        for (int l = 0; l < 4; ++l)
            for (int j = 0; j < 3; ++j)
                x[12 + 3 * l + j] = (c[l] > 0) ? g->foot[12 * (size_t)(w.k0 + k) + 3 * l + j] : g->qJ[12 * (size_t)(w.k0 + k) + 3 * l + j];
    };
    size_t node = 0;
    for (int i = 0; i < n_phases; ++i) {
        const float t_offset = out->start_time[i] - out->start_time[0];  // set_time_offset, HKDProblem.cpp:98
        for (int k = 0; k <= out->horizon[i]; ++k, ++node) {
            // time seen by the cost callbacks: float(t_offset + k*dt) with dt widened to double (SinglePhase.cpp:243,254)
            const float tc = (float)((double)t_offset + k * out->dt);
            const int kc = w.index_at(tc);
            fill_state_ref(kc, out->xr + 24 * node);
            for (int j = 0; j < 12; ++j) out->ur[24 * node + j] = g->grf[12 * (size_t)(w.k0 + kc) + j];  // qJd never loaded: zeros
            const double* bs = &g->body_state[12 * (size_t)(w.k0 + kc)];
            for (int l = 0; l < 4; ++l)
                for (int j = 0; j < 3; ++j)
                    out->prel_r[12 * node + 3 * l + j] = g->foot[12 * (size_t)(w.k0 + kc) + 3 * l + j] - bs[3 + j];
            // time used for the initial guess: float(phase_start + k*dt_sim), all float (HKDProblem.cpp:86-90)
            const float ti = out->start_time[i] + (float)k * dt_sim;
            fill_state_ref(w.index_at(ti), out->xinit + 24 * node);
        }
    }
    return HSDDP_OK;
}

void hkd_schedule_free(hsddp_schedule* s) {
    if (!s) return;
    std::free(s->xr); std::free(s->ur); std::free(s->prel_r); std::free(s->xinit);
    s->xr = s->ur = s->prel_r = s->xinit = nullptr;
}

void hkd_compute_state(const double eul[3], const double pos[3], const double qJ[12], const int32_t contact[4], double qdummy[12]) {
    hkd::hkd_state(eul, pos, qJ, mask_of(contact), qdummy);
}

void hkd_default_x0(const hsddp_schedule* s, double x0[24]) {
    const double body[12] = {0, 0, 0, 0, 0, 0.2486, 0, 0, 0, 0, 0, 0};
    const double qJ[12] = {0, -0.8, 1.6, 0, -0.8, 1.6, 0, -0.8, 1.6, 0, -0.8, 1.6};
    for (int i = 0; i < 12; ++i) x0[i] = body[i];
    hkd::hkd_state(body, body + 3, qJ, mask_of(s->contact[0]), x0 + 12);
}

void hkd_model_dynamics(const double x[24], const double u[24], double dt, const int32_t contact[4], double xnext[24]) {
    hkd::dynamics(x, u, dt, mask_of(contact), xnext);
}

void hkd_model_dynamics_partial(const double x[24], const double u[24], double dt, const int32_t contact[4], double A[576], double B[576]) {
    double R40[hkd::kCrSize] = {0};
    const unsigned m = mask_of(contact);
    hkd::dynamics_partial_record(x, u, dt, m, R40);
    hkd::expand_AB(R40, dt, m, A, B);
}

void hkd_model_foot_position(const double pos[3], const double eul[3], const double qleg[3], int leg, double p[3]) {
    hkd::foot_position(pos, eul, qleg, leg, p);
}

void hkd_model_foot_jacobian(const double pos[3], const double eul[3], const double qleg[3], int leg, double J[54]) {
    (void)pos;
    double Jc[18];
    hkd::foot_jacobian_compact(eul, qleg, leg, Jc);
    for (int i = 0; i < 54; ++i) J[i] = 0.0;
    for (int i = 0; i < 3; ++i) {
        J[i + 3 * i] = 1.0;
        for (int c = 0; c < 3; ++c) {
            J[i + 3 * (3 + c)] = Jc[i * 6 + c];
            J[i + 3 * (6 + 3 * leg + c)] = Jc[i * 6 + 3 + c];
        }
    }
}

}  // extern "C"
