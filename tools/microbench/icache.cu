// Does instruction-cache pressure explain the slow in-kernel Gauss-Jordan?  One warp alternates between
// gauss_jordan12 and a block of straight-line filler code of configurable size (KB of SASS).
#include <cstdio>
#include "../../hkd-mpc_b200/csrc/hsddp_sweep.cuh"
using namespace hsddp;
template <int FILL>  // FILL x 64 dependent-free FMAs (each DFMA = 16 B of code)
__global__ void k_ic(double* out, long long* cyc, const double* in, int reps) {
    __shared__ __align__(16) double sbuf[128];
    __shared__ __align__(16) double Q[24 * 24];
    for (int e = threadIdx.x; e < 576; e += blockDim.x) Q[e] = in[e];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0;
    double f0 = in[lane], f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
    long long tgj = 0, tfill = 0;
    for (int it = 0; it < reps; ++it) {
        double col[12];
#pragma unroll
        for (int r = 0; r < 12; ++r) col[r] = Q[r * 24 + (lane % 24)] + ((r == lane % 12 && lane < 12) ? 30.0 + it : 0.0);
        __syncwarp();
        long long t0 = clock64();
        bool ok = gauss_jordan12(col, sbuf + 32 * warp);
        long long t1 = clock64();
        tgj += t1 - t0;
#pragma unroll
        for (int r = 0; r < 12; ++r) acc += col[r];
        acc += ok;
#pragma unroll
        for (int i = 0; i < FILL * 16; ++i) {  // 4 independent chains, straight-line
            f0 = fma(f0, 1.0000001, 1e-9 * i); f1 = fma(f1, 0.9999999, 2e-9 * i); f2 = fma(f2, 1.0000002, 3e-9); f3 = fma(f3, 0.9999998, 4e-9);
        }
        long long t2 = clock64();
        tfill += t2 - t1;
    }
    if (threadIdx.x == 0) { cyc[0] = tgj / reps; cyc[1] = tfill / reps; }
    out[threadIdx.x] = acc + f0 + f1 + f2 + f3;
}
template <int FILL> void run(double* out, long long* cyc, double* in, int threads) {
    for (int rep = 0; rep < 2; ++rep) { k_ic<FILL><<<1, threads>>>(out, cyc, in, 50); cudaDeviceSynchronize(); }
    printf("filler %4d DFMA (%5.1f KB code)  threads %3d : gauss_jordan12 %lld cyc/call, filler %lld cyc (%.2f cyc/instr)  %s\n", FILL * 64, FILL * 64 * 16 / 1024.0,
           threads, cyc[0], cyc[1], (double)cyc[1] / (FILL * 64 + 1), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    double *in, *out; long long* cyc;
    cudaMallocManaged(&in, 576 * 8); cudaMallocManaged(&out, 128 * 8); cudaMallocManaged(&cyc, 8 * 8);
    for (int i = 0; i < 576; ++i) in[i] = 0.01 * ((i * 7) % 13) + ((i / 24 == i % 24) ? 1.0 : 0.0);
    run<1>(out, cyc, in, 32);
    run<8>(out, cyc, in, 32);
    run<24>(out, cyc, in, 32);
    run<48>(out, cyc, in, 32);
    run<96>(out, cyc, in, 32);
    run<24>(out, cyc, in, 128);
    run<96>(out, cyc, in, 128);
    return 0;
}
