// Latency micro-benchmarks on B200 for the building blocks of the Riccati stage (one warp, clock64).
#include <cstdio>
#include "../../hkd-mpc_b200/csrc/hsddp_sweep.cuh"
using namespace hsddp;
__global__ void k_lat(double* out, long long* cyc, const double* in) {
    __shared__ __align__(16) double sbuf[128];
    __shared__ __align__(16) double Q[24 * 24];
    for (int e = threadIdx.x; e < 576; e += blockDim.x) Q[e] = in[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double acc = 0;
    long long t0, t1;
    // 1. dependent DFMA chain
    double x = in[lane];
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = fma(x, 1.0000001, 1e-9);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 64;
    acc += x;
    // 2. dependent DMMA chain
    double c[2] = {0, 0};
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) dmma884(c, x, 1e-3);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = (t1 - t0) / 32;
    acc += c[0] + c[1];
    // 3. independent DMMAs (4 accumulators)
    double c4[4][2] = {{0,0},{0,0},{0,0},{0,0}};
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) { dmma884(c4[0], x, 1e-3); dmma884(c4[1], x, 2e-3); dmma884(c4[2], x, 3e-3); dmma884(c4[3], x, 4e-3); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = (t1 - t0) / 64;
    acc += c4[0][0] + c4[1][1] + c4[2][0] + c4[3][1];
    // 4. dependent division chain
    double d = in[lane] + 2.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) d = 1.0 / d + 1.5;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = (t1 - t0) / 16;
    acc += d;
    // 5. dependent LDS chain (pointer chase through smem)
    int idx = lane;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) idx = (int)Q[idx & 511] & 511;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = (t1 - t0) / 32;
    acc += idx;
    // 6. Gauss-Jordan 12 (2x2 pivots) on an SPD tableau
    double col[12];
#pragma unroll
    for (int r = 0; r < 12; ++r) col[r] = Q[r * 24 + (lane % 24)] + ((r == lane % 12 && lane < 12) ? 30.0 : 0.0);
    __syncwarp();
    t0 = clock64();
    bool ok = gauss_jordan12(col, sbuf);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = (t1 - t0);
#pragma unroll
    for (int r = 0; r < 12; ++r) acc += col[r];
    acc += ok;
    // 7. shfl chain
    double sh = acc;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) sh = __shfl_sync(0xffffffffu, sh, (lane + 1) & 31) + 1.0;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = (t1 - t0) / 32;
    acc += sh;
    // 8. __syncthreads cost (whole block)
    __syncthreads();
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = (t1 - t0) / 16;
    out[threadIdx.x] = acc;
}
int main() {
    double *in, *out; long long* cyc;
    cudaMallocManaged(&in, 576 * 8); cudaMallocManaged(&out, 128 * 8); cudaMallocManaged(&cyc, 8 * 8);
    for (int i = 0; i < 576; ++i) in[i] = 0.01 * ((i * 7) % 13) + ((i / 24 == i % 24) ? 1.0 : 0.0);
    for (int rep = 0; rep < 3; ++rep) { k_lat<<<1, 128>>>(out, cyc, in); cudaDeviceSynchronize(); }
    const char* nm[8] = {"DFMA dependent", "DMMA dependent", "DMMA 4-way independent (per op)", "1/x + add dependent", "LDS dependent", "gauss_jordan12 total", "SHFL+DADD dependent", "__syncthreads (128 thr)"};
    for (int i = 0; i < 8; ++i) printf("%-36s %lld cycles\n", nm[i], cyc[i]);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
