// ORACLE (test infrastructure only — never linked by the product).
//
// CPU restatement, for the <double,24,24,0> HKD instantiation, of the
// reference's Hybrid-Systems DDP hot path:
//   MultiPhaseDDP<T>          HSDDPSolver/source/MultiPhaseDDP.cpp:20-541
//   SinglePhase<T,24,24,0>    HSDDPSolver/source/SinglePhase.cpp:145-426
//   Trajectory<T,24,24,0>     HSDDPSolver/source/TrajectoryManagement.cpp:5-35,110-115,210-238
//   cost container + costs    HSDDPSolver/source/SinglePhaseInterface.cpp:55-165,
//                             HKDMPC/HKD-TrajOpt/HKDCost.h:11-73, HKDCost.cpp:5-66
//   ReB / AL constraints      HSDDPSolver/header/ConstraintsBase.h:168-263,349-399,
//                             HKDMPC/HKD-TrajOpt/HKDConstraints.cpp:7-171
//   reset map                 HKDMPC/HKD-TrajOpt/HKDReset.h:41-136
//   reference lookup          Reference/QuadReference.cpp:6-26,65-100, HKDReference.cpp:8-57
//   problem assembly          HKDMPC/HKD-TrajOpt/HKDProblem.cpp:15-111,225-310
//   options                   HSDDPSolver/common/HSDDP_CompoundTypes.h:18-60
// including the behaviours listed in SURVEY.md §9 (Q1-Q19).
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or recorded
// outputs for this path and cannot be built here (Eigen/Boost/LCM absent), so
// this restatement is pinned only (a) at the model level, against the
// reference's own CasADi C compiled unmodified (oracle/_ref), and (b) by
// internal invariants (tests/test_oracle_*.py).
#pragma once
#include <cstdint>
#include <vector>
#include "hkd_model.hpp"
#include "linalg.hpp"

namespace oracle {

// HSDDP_OPTION as actually consumed by solve() (Q4): values of
// HKDMPC/settings/ddp_setting.info, with update_regularization left at its
// struct default 2 because loadHSDDPSetting never reads it.
struct Options {
    double alpha = 0.1;
    double gamma = 0.01;
    double update_penalty = 5;
    double update_relax = 1;
    double update_regularization = 2;
    double update_ReB = 1;
    int max_DDP_iter = 10;
    int max_AL_iter = 5;
    double cost_thresh = 1e-3;
    double tconstr_thresh = 1e-3;
    double pconstr_thresh = 1e-3;
    double dynamics_feas_thresh = 1e-3;
    double merit_scale = 0.2;
    double merit_offset = 1e2;
    int AL_active = 1;
    int ReB_active = 1;
    int MS = 1;
};

// HKDMPC/settings/constraint_params.info
struct ConstraintParams {
    double grf_delta = 0.1, grf_delta_min = 0.1, grf_eps = 0.1;
    double td_sigma = 50, td_sigma_max = 1e4, td_lambda = 0;
    double mu = 0.7;  // HKDConstraints.h:17
};

// What QuadReference::load_top_level_data leaves in tp_data: every value went
// through std::stof, so it is exactly representable in float (Q11).
struct GaitTable {
    int n = 0;
    float dt = 0.f;
    std::vector<double> body_state;       // n x 12
    std::vector<double> qJ;               // n x 12
    std::vector<double> foot_placements;  // n x 12
    std::vector<double> grf;              // n x 12
    std::vector<int> contact;             // n x 4
};

// QuadReference restricted to what the solve touches.
struct QuadReference {
    const GaitTable* tp = nullptr;
    int k0 = 0;   // window start inside the top-level table (k_cur of the reference, offset by the initial window start)
    int sz = 0;   // round(plan/dt)+1 ; the window holds sz+1 samples
    float dt = 0.f;
    float t_cur = 0.f, dur = 0.f, start_time = 0.f, end_time = 0.f;  // QuadReference.cpp:6-26,33-47
    void initialize(const GaitTable* table, int window_start, float plan_horizon);
    void step(float dt_sim);                       // QuadReference.cpp:33-47: shift the window by dt_sim
    int index_at_t(float t) const;                 // QuadReference.cpp:65-80 (nearest sample, float arithmetic)
    const int* contact_at_t(float t) const;        // :86-100
    const double* body_state(int k) const { return &tp->body_state[12 * (size_t)(k0 + k)]; }
    const double* qJ(int k) const { return &tp->qJ[12 * (size_t)(k0 + k)]; }
    const double* foot(int k) const { return &tp->foot_placements[12 * (size_t)(k0 + k)]; }
    const double* grf(int k) const { return &tp->grf[12 * (size_t)(k0 + k)]; }
    const int* contact(int k) const { return &tp->contact[4 * (size_t)(k0 + k)]; }
};

struct RCost {  // RCostData<T,24,24,0>, HSDDP_CompoundTypes.h:90-127
    double l;
    Vec24 lx, lu;
    Mat24 lxx, lux, luu;
    void zero() { l = 0; lx.zero(); lu.zero(); lxx.zero(); lux.zero(); luu.zero(); }
};
struct TCost {  // TCostData<T,24>, :130-150
    double Phi;
    Vec24 Phix;
    Mat24 Phixx;
    void zero() { Phi = 0; Phix.zero(); Phixx.zero(); }
};

struct RebParam { double delta, delta_min, eps; };
struct AlParam { double lambda, sigma, sigma_max; };

// One contact phase = SinglePhase<double,24,24,0> + its Trajectory + its
// cost/constraint objects, flattened.
struct Phase {
    int horizon = 0;
    int contact[4] = {0, 0, 0, 0};
    int next_contact[4] = {0, 0, 0, 0};
    float start_time = 0.f;  // pdata->phase_start_times[i]
    float t_offset = 0.f;    // start_time - start_time[0]
    double dt = 0.0;         // (double)(float)0.01

    // GRF path constraint (present iff any stance leg): 5 rows per stance leg
    int n_stance = 0;
    int stance_legs[4];
    int n_path = 0;                       // 5*n_stance
    std::vector<double> g;                // horizon x n_path
    std::vector<RebParam> reb;            // horizon x n_path
    double path_max_violation = 0.0;
    // touchdown terminal constraints (legs going 0 -> 1).  One TouchDownConstraint OBJECT per add_tconstr_one_phase call
    // that found a touchdown leg: HKDProblem::update calls it again when a phase reaches its end, so a phase can carry
    // two objects on the same legs, each with its own AL parameters (HKDProblem.cpp:205-208,268-310)
    struct TdSet {
        int n_td = 0;
        int td_legs[4] = {0, 0, 0, 0};
        double h[4] = {0, 0, 0, 0};
        Vec24 hx[4];
        AlParam al[4];
        double max_violation = 0.0;
    };
    std::vector<TdSet> tds;
    bool has_tconstr = false;   // add_tconstr_one_phase has run: next_contact / reset map are bound
    bool reach_end = false;     // pdata->is_phase_reach_end
    float end_time = 0.f;       // pdata->phase_end_times[i]
    int ss_size = 0;            // SS_set = {0 .. ss_size-1} (update_SS_config); 0 = empty (SinglePhase::initialization)
    int n_td_total() const { int n = 0; for (auto& t : tds) n += t.n_td; return n; }

    // Trajectory
    std::vector<Vec24> Xbar, X, Xsim, Defect, Defect_bar, dX, G;  // horizon+1
    std::vector<Vec24> Ubar, U, dU;                                // horizon
    std::vector<Mat24> A, B, H, K;                                 // A,H,K: horizon+1 ; B: horizon
    std::vector<RCost> rcost;                                      // horizon
    TCost tcost;

    // SinglePhase scratch / state
    Vec24 x_init, dx_init;
    double actual_cost = 0, dV_1 = 0, dV_2 = 0;
};

// One record per DDP iteration (inner loop body), in execution order.
struct IterRecord {
    int outer, inner;
    double cost_before, feas_before;   // compute_cost / feas at the top of the iteration
    double reg_used;                   // regularisation of the successful backward sweep
    int n_sweeps;                      // backward sweeps incl. retries
    double dV_1, dV_2, merit_rho;
    double eps_accepted;               // 0 = all step sizes rejected, -1 = early exit before line search
    int n_trials;
    double cost_after, feas_after, max_tconstr, max_pconstr;
};

struct SolveResult {
    int status = 0;        // 0 constraints satisfied, 1 constraints stalled, 2 max AL iterations, 3 regularisation overflow
    int n_iter = 0;        // total DDP iterations
    int n_outer = 0;
    int n_sweeps = 0;
    double cost = 0, feas = 0, max_tconstr = 0, max_pconstr = 0;
    double cost0 = 0, feas0 = 0;
    std::vector<IterRecord> trace;
    // get_solver_info buffers (MultiPhaseDDP.cpp:277-280,368-371,532-541)
    std::vector<float> cost_buffer, dyn_feas_buffer, eqn_feas_buffer, ineq_feas_buffer;
};

struct Problem {
    Model model;
    ConstraintParams cparams;
    QuadReference ref;
    float plan_duration = 0.6f;
    float dt_sim = 0.01f;
    float dt_mpc = 0.01f;
    std::vector<Phase> phases;
    Vec24 x0;

    // ---- HKDProblem::initialization (HKDProblem.cpp:15-111) ----
    void build(const GaitTable* table, int window_start, float plan_dur, ModelKind kind, const ConstraintParams& cp);
    // ---- HKDProblem::update (HKDProblem.cpp:117-222): receding-horizon shift by one MPC step ----
    void update();
    void create_phase(Phase& ph, int horizon);      // Trajectory(dt, horizon) + create_problem_one_phase (:225-265) + SinglePhase::initialization
    void add_tconstr(int idx);                      // add_tconstr_one_phase (:268-310)
    void phase_pop_front(Phase& ph);                // SinglePhase::pop_front (SinglePhase.cpp:496-501)
    void phase_push_back_default(Phase& ph);        // SinglePhase::push_back_default (:486-491)
    // default initial condition of HKDMPC.cpp:44-54
    void default_x0(Vec24& x) const;

    // ---- MultiPhaseDDP public step API (MultiPhaseDDP.h:42-69) ----
    void linear_rollout(double eps, const Options& opt);
    bool hybrid_rollout(double eps, const Options& opt);
    bool line_search(const Options& opt, int* n_trials, double* eps_out);
    void compute_cost(const Options& opt);
    void LQ_approximation(const Options& opt);
    bool backward_sweep(double regularization);
    bool backward_sweep_regularized(double& regularization, const Options& opt, int* n_sweeps);
    void update_nominal_trajectory();
    void update_AL_params(const Options& opt);
    void update_REB_params(const Options& opt);
    double measure_dynamics_feasibility();
    void solve(const Options& opt, SolveResult& out);

    // solver scalars (MultiPhaseDDP.h:92-104)
    double actual_cost = 0, merit = 0, feas = 0, dV_1 = 0, dV_2 = 0;
    double max_tconstr_prev = 0, max_pconstr_prev = 0, max_tconstr = 0, max_pconstr = 0, merit_rho = 0;

    // ---- per-phase pieces (SinglePhase / callbacks) ----
    void reference_at_t(float t, double xr[24], double ur[24], int* sample_idx) const;  // HKDReference.cpp:8-57
    void resetmap(const Phase& ph, const Vec24& x, Vec24& xnext) const;                 // HKDReset.h:41-75
    void resetmap_partial(const Phase& ph, const Vec24& x, Mat24& Px) const;            // HKDReset.h:78-136
    bool phase_hybrid_rollout(Phase& ph, double eps, const Options& opt);
    void phase_linear_rollout(Phase& ph, double eps);
    void phase_compute_cost(Phase& ph, const Options& opt);
    void phase_LQ_approximation(Phase& ph, const Options& opt);
    bool phase_backward_sweep(Phase& ph, double reg, const Vec24& Gprime, const Mat24& Hprime);
    void running_cost(const Phase& ph, int k, RCost& rc) const;
    void running_cost_par(const Phase& ph, int k, RCost& rc) const;
    void terminal_cost(const Phase& ph, TCost& tc) const;
    void terminal_cost_par(const Phase& ph, TCost& tc) const;
    void grf_violation(Phase& ph, int k);
    void td_violation(Phase& ph);
    void td_partial(Phase& ph);
};

}  // namespace oracle
