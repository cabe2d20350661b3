// Gauss-Jordan (P3) latency under concurrency: 1..4 warps of a block, 1..6 blocks per SM.
#include <cstdio>
#include "../../hkd-mpc_b200/csrc/hsddp_sweep.cuh"
using namespace hsddp;
__global__ void __launch_bounds__(128, 6) k_gj(double* out, long long* cyc, const double* in, int nw, int reps) {
    __shared__ __align__(16) double sbuf[128];
    __shared__ __align__(16) double Q[24 * 24];
    for (int e = threadIdx.x; e < 576; e += blockDim.x) Q[e] = in[e];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0;
    if (warp < nw) {
        long long t0 = clock64();
        for (int it = 0; it < reps; ++it) {
            double col[12];
#pragma unroll
            for (int r = 0; r < 12; ++r) col[r] = Q[r * 24 + (lane % 24)] + ((r == lane % 12 && lane < 12) ? 30.0 : 0.0) + acc * 1e-30;
            __syncwarp();
            bool ok = gauss_jordan12(col, sbuf + 32 * warp);
#pragma unroll
            for (int r = 0; r < 12; ++r) acc += col[r];
            acc += ok;
        }
        long long t1 = clock64();
        if (lane == 0) cyc[blockIdx.x * 4 + warp] = (t1 - t0) / reps;
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}
int main() {
    double *in, *out; long long* cyc;
    cudaMallocManaged(&in, 576 * 8); cudaMallocManaged(&out, 148 * 6 * 128 * 8); cudaMallocManaged(&cyc, 148 * 6 * 4 * 8);
    for (int i = 0; i < 576; ++i) in[i] = 0.01 * ((i * 7) % 13) + ((i / 24 == i % 24) ? 1.0 : 0.0);
    for (int bps = 1; bps <= 6; bps += 5)
        for (int nw = 1; nw <= 4; ++nw) {
            for (int rep = 0; rep < 2; ++rep) { k_gj<<<148 * bps, 128>>>(out, cyc, in, nw, 64); cudaDeviceSynchronize(); }
            long long s = 0; int n = 0;
            for (int b = 0; b < 148 * bps; ++b) for (int w = 0; w < nw; ++w) { s += cyc[b * 4 + w]; ++n; }
            printf("blocks/SM %d  GJ warps/block %d : %lld cycles per gauss_jordan12 (+tableau load)\n", bps, nw, s / n);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
