// Minimal C++ use of the drop-in surface: the equivalent of HKDMPCSolver::initialize
// (HKDMPC/HKDMPC.cpp:20-83) for a batch of problems, through hkd-mpc_b200/host/MultiPhaseDDP.hpp.
//   g++ -std=c++17 examples/solve_trot.cpp -Lhkd-mpc_b200 -lhsddp_b200 -Wl,-rpath,hkd-mpc_b200 -o solve_trot
//   ./solve_trot <quad_reference.csv> [n_problems]
#include <cstdio>
#include <cstdlib>
#include "../hkd-mpc_b200/host/MultiPhaseDDP.hpp"

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s quad_reference.csv [n]\n", argv[0]); return 2; }
    const int n = argc > 2 ? std::atoi(argv[2]) : 4;
    try {
        hsddp_b200::QuadReference quad_reference;
        quad_reference.load_top_level_data(argv[1]);
        hsddp_b200::Schedule schedule(quad_reference, 0, 0.6f);  // plan_duration = .6, timeStep = .01 (HKDMPC.cpp:26-28)
        std::vector<double> x0;
        for (int i = 0; i < n; ++i) {
            std::vector<double> xi = schedule.default_x0();
            xi[3] += 0.001 * i;  // shift the body a little for every problem
            x0.insert(x0.end(), xi.begin(), xi.end());
        }
        hsddp_b200::MultiPhaseDDP<double> solver(0);
        solver.set_multiPhaseProblem({&schedule}, std::vector<int32_t>(n, 0));
        solver.set_initial_condition(x0);
        hsddp_b200::HSDDP_OPTION ddp_options;
        solver.solve(ddp_options);
        for (const hsddp_info& info : solver.get_info())
            std::printf("status %d  iterations %d  total cost = %.8f  dynamics infeasibility = %.3e\n", info.status, info.n_iter, info.cost, info.feas);
        // the phase interface of the reference's problem assembly compiles against the shim (its plug-ins are accepted and ignored)
        hsddp_b200::SinglePhase<double, 24, 24, 0> phase;
        phase.set_dynamics([](double*, double*, double*, double*, double) {});
        phase.set_time_offset(0.f);
        phase.update_SS_config(12);
        const std::vector<double> feas = solver.measure_dynamics_feasibility();
        if (feas.size() != (size_t)n) return 3;
        // the same batch sharded over "two GPUs" (device 0 twice: the logic is what is checked here)
        hsddp_b200::MultiGpuDDP multi({0, 0});
        multi.set_multiPhaseProblem({&schedule}, std::vector<int32_t>(n, 0));
        multi.set_initial_condition(x0);
        multi.solve(ddp_options);
        const std::vector<hsddp_info> a = solver.get_info(), b2 = multi.get_info();
        for (int i = 0; i < n; ++i)
            if (a[i].n_iter != b2[i].n_iter || a[i].status != b2[i].status || a[i].cost != b2[i].cost) { std::fprintf(stderr, "multi-GPU shard %d differs\n", i); return 4; }
        // what the reference publishes on "mpc_command" (HKDMPC.cpp:243-298)
        const std::vector<hsddp_mpc_command> cmd = solver.get_mpc_command(8);
        std::printf("command of problem 0: %d steps, first GRF z of leg 0 = %.4f N, feedback[0][2][5] = %.4f\n", cmd[0].N_mpcsteps,
                    cmd[0].hkd_controls[0][2], cmd[0].feedback[0][2][5]);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
