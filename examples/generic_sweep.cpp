// The reference's other instantiations through the C++ shim: SinglePhase<double,36,12,12>::backward_sweep and
// linear_rollout (HSDDPSolver/source/SinglePhase.cpp:145-178, 299-367) on plug-in outputs read from a binary file.
//
//   generic_sweep <file> <n_problems> <horizon> <regularization> <eps>
//
// <file>: the 14 input arrays of include/hsddp_b200.h (HSDDP_PH_A .. HSDDP_PH_DEFECT), doubles, column-major matrices,
// problem-major, one after the other.  Prints per problem: the reference's bool, dV_1, dV_2 of the sweep, K(0,0) of stage 0,
// dV_1 of the linear rollout and dX of the last node, component 0.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../hkd-mpc_b200/host/MultiPhaseDDP.hpp"

int main(int argc, char** argv) {
    if (argc < 6) { std::fprintf(stderr, "usage: %s file n horizon reg eps\n", argv[0]); return 2; }
    constexpr size_t xs = 36, us = 12, ys = 12;
    const int n = std::atoi(argv[2]), N = std::atoi(argv[3]);
    const double reg = std::atof(argv[4]), eps = std::atof(argv[5]);
    const size_t counts[HSDDP_PH_N_INPUTS] = {N * xs * xs, N * xs * us, N * ys * xs, N * ys * us, N * xs, N * us, N * ys, N * xs * xs, N * us * us,
                                              N * us * xs, N * ys * ys, xs, xs * xs, (N + 1) * xs};
    std::FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 1; }
    try {
        hsddp_b200::SinglePhaseSweeps<double, xs, us, ys> phase(N, n);
        for (int w = 0; w < HSDDP_PH_N_INPUTS; ++w) {
            std::vector<double> v(counts[w] * n);
            if (std::fread(v.data(), sizeof(double), v.size(), f) != v.size()) { std::fprintf(stderr, "short read\n"); return 1; }
            phase.set(w, v);
        }
        std::fclose(f);
        const std::vector<int32_t> ok = phase.backward_sweep(reg);
        std::vector<double> dV(2 * n), K((size_t)n * N * us * xs), dVl(2 * n), dX((size_t)n * (N + 1) * xs);
        phase.get(HSDDP_PH_OUT_DV, dV);
        phase.get(HSDDP_PH_OUT_K, K);
        phase.linear_rollout(eps);
        phase.get(HSDDP_PH_OUT_DV, dVl);
        phase.get(HSDDP_PH_OUT_DX, dX);
        for (int i = 0; i < n; ++i)
            std::printf("ok=%d dV_1=%.17g dV_2=%.17g K00=%.17g lr_dV_1=%.17g dXN0=%.17g\n", ok[i], dV[2 * i], dV[2 * i + 1], K[(size_t)i * N * us * xs],
                        dVl[2 * i], dX[((size_t)i * (N + 1) + N) * xs]);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
