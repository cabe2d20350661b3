// Latency and accuracy of the 2x2-block vs 3x3-block Gauss-Jordan (one warp; and 6 blocks x 2 warps per SM).
#include <cstdio>
#include <cmath>
#include "../../hkd-mpc_b200/csrc/hsddp_sweep.cuh"
using namespace hsddp;
template <int KIND>
__global__ void __launch_bounds__(128, 6) k_gj(double* out, long long* cyc, const double* in, int nw, int reps) {
    __shared__ __align__(16) double sbuf[160];
    __shared__ __align__(16) double Q[24 * 24];
    for (int e = threadIdx.x; e < 576; e += blockDim.x) Q[e] = in[e];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0;
    double col[12];
    if (warp < nw) {
        long long t0 = clock64();
        for (int it = 0; it < reps; ++it) {
#pragma unroll
            for (int r = 0; r < 12; ++r) col[r] = Q[r * 24 + (lane % 24)] + acc * 1e-30;
            __syncwarp();
            bool ok = KIND == 0 ? gauss_jordan12(col, sbuf + 40 * warp) : gauss_jordan12_b3(col, sbuf + 40 * warp);
#pragma unroll
            for (int r = 0; r < 12; ++r) acc += col[r];
            acc += ok;
        }
        long long t1 = clock64();
        if (lane == 0) cyc[blockIdx.x * 4 + warp] = (t1 - t0) / reps;
    }
    if (blockIdx.x == 0 && warp == 0) for (int r = 0; r < 12; ++r) out[(KIND * 32 + lane) * 12 + r] = col[r];
}
int main() {
    double *in, *out; long long* cyc;
    cudaMallocManaged(&in, 576 * 8); cudaMallocManaged(&out, 2 * 32 * 12 * 8); cudaMallocManaged(&cyc, 148 * 6 * 4 * 8);
    // SPD 12x12 block in the top-left corner (columns 0..11), right-hand sides in columns 12..23
    for (int i = 0; i < 24; ++i) for (int j = 0; j < 24; ++j) in[i * 24 + j] = 0.01 * (((i * 7 + j * 3) % 13) - 6);
    for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) { double s = 0; for (int k = 0; k < 12; ++k) s += 0.05 * (((i * 5 + k * 3) % 7) - 3) * 0.05 * (((j * 5 + k * 3) % 7) - 3); in[i * 24 + j] = s + (i == j ? 0.02 : 0.0); }
    for (int bps = 1; bps <= 6; bps += 5)
        for (int kind = 0; kind < 2; ++kind) {
            for (int rep = 0; rep < 2; ++rep) {
                if (kind == 0) k_gj<0><<<148 * bps, 128>>>(out, cyc, in, 2, 64); else k_gj<1><<<148 * bps, 128>>>(out, cyc, in, 2, 64);
                cudaDeviceSynchronize();
            }
            long long s = 0; int n = 0;
            for (int b = 0; b < 148 * bps; ++b) for (int w = 0; w < 2; ++w) { s += cyc[b * 4 + w]; ++n; }
            printf("blocks/SM %d  %s : %lld cycles per elimination (+tableau load)\n", bps, kind ? "3x3 blocks (4 steps)" : "2x2 blocks (6 steps)", s / n);
        }
    double md = 0, mx = 0;
    for (int i = 0; i < 32 * 12; ++i) { md = fmax(md, fabs(out[i] - out[32 * 12 + i])); mx = fmax(mx, fabs(out[i])); }
    printf("max |x_2x2 - x_3x3| = %.3e  (max |x| = %.3e)\n%s\n", md, mx, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
