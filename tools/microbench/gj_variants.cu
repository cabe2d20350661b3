// Latency study of the 12x12 tableau elimination (P3 of the Riccati stage): which restructuring shortens the serial chain?
//   0  gauss_jordan12            (2x2 pivots, rolled, IEEE 1/det in the chain)
//   1  same, branch-free Newton reciprocal
//   2  gauss_jordan12_la         (lookahead pivot block, rolled, IEEE 1/det)
//   3  lookahead, loop rotated so that the reciprocal is issued before the elimination, Newton reciprocal, double-buffered publish
//   4  variant 3 fully unrolled
//   5  variant 1 fully unrolled
//   6  4x4 pivot blocks (three steps), Newton reciprocals   [added after the last GPU run of round 1: not yet measured]
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gj_variants gj_variants.cu
#include <cstdio>
#include <cmath>
#include "../../hkd-mpc_b200/csrc/hsddp_sweep.cuh"
using namespace hsddp;

__device__ __forceinline__ double newton_rcp(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
}

// (kept here for the record; it is slower)  The same elimination with the pivot inversion taken OFF the critical path.  In gauss_jordan12 every step is one serial
// chain: publish the pivot columns -> read the 2x2 pivot block -> det -> 1/det (71 cycles) -> multipliers -> eliminate.
// Here the two columns that pivot NEXT also publish their rows 0..3, and every lane applies the current step to those
// four numbers on the side: that yields the next pivot block one step early, so its determinant and reciprocal are
// computed while the current elimination, the next publish and the next read are in flight.  The chain of a step shrinks
// to publish -> read -> multipliers -> eliminate.  `sbuf`: 32 doubles per warp.  Same pivots, same verdict.
__device__ __forceinline__ bool gauss_jordan12_la(double (&v)[12], double* sbuf) {
    const int lane = threadIdx.x & 31;
    bool ok = true;
    double* look = sbuf + 24;  // [2][4]: rows 0..3 of the two columns that pivot next
    if (lane < 2) *reinterpret_cast<double2*>(look + 4 * lane) = make_double2(v[0], v[1]);
    __syncwarp();
    double pa, pb, pc, pd, rdet;  // pivot block [pa pb; pc pd] of the coming step and 1 / det
    {
        const double2 q0 = *reinterpret_cast<const double2*>(look), q1 = *reinterpret_cast<const double2*>(look + 4);
        pa = q0.x; pc = q0.y; pb = q1.x; pd = q1.y;
        const double det = pa * pd - pb * pc;
        if (pa < 0.0 || det < 0.0) ok = false;
        rdet = 1.0 / det;
    }
    __syncwarp();
#pragma unroll 1
    for (int step = 0; step < 6; ++step) {
        const int role = (lane >> 1) - step;  // 0: pivot column of this step, 1: pivot column of the next step
        if (role == 0) {
            double2* dst = reinterpret_cast<double2*>(sbuf + 12 * (lane & 1));
#pragma unroll
            for (int r = 0; r < 12; r += 2) dst[r >> 1] = make_double2(v[r], v[r + 1]);
        } else if (role == 1) {
            double2* dst = reinterpret_cast<double2*>(look + 4 * (lane & 1));
            dst[0] = make_double2(v[0], v[1]);
            dst[1] = make_double2(v[2], v[3]);
        }
        __syncwarp();
        const double2 a2 = *reinterpret_cast<const double2*>(sbuf + 2);
        const double2 b2 = *reinterpret_cast<const double2*>(sbuf + 14);
        // side computation: rows 2,3 of the next two pivot columns after this step = the next pivot block
        double na, nb, nc, nd;
        {
            const double2 n00 = *reinterpret_cast<const double2*>(look), n01 = *reinterpret_cast<const double2*>(look + 2);
            const double2 n10 = *reinterpret_cast<const double2*>(look + 4), n11 = *reinterpret_cast<const double2*>(look + 6);
            const double u0 = (pd * n00.x - pb * n00.y) * rdet, u1 = (pa * n00.y - pc * n00.x) * rdet;
            const double w0 = (pd * n10.x - pb * n10.y) * rdet, w1 = (pa * n10.y - pc * n10.x) * rdet;
            na = fma(-b2.x, u1, fma(-a2.x, u0, n01.x));
            nc = fma(-b2.y, u1, fma(-a2.y, u0, n01.y));
            nb = fma(-b2.x, w1, fma(-a2.x, w0, n11.x));
            nd = fma(-b2.y, w1, fma(-a2.y, w0, n11.y));
        }
        const double t0 = (pd * v[0] - pb * v[1]) * rdet;
        const double t1 = (pa * v[1] - pc * v[0]) * rdet;
        v[0] = fma(-b2.x, t1, fma(-a2.x, t0, v[2]));  // eliminate and rotate in one go
        v[1] = fma(-b2.y, t1, fma(-a2.y, t0, v[3]));
#pragma unroll
        for (int r = 4; r < 12; r += 2) {
            const double2 a = *reinterpret_cast<const double2*>(sbuf + r);
            const double2 b = *reinterpret_cast<const double2*>(sbuf + 12 + r);
            v[r - 2] = fma(-b.x, t1, fma(-a.x, t0, v[r]));
            v[r - 1] = fma(-b.y, t1, fma(-a.y, t0, v[r + 1]));
        }
        v[10] = t0;
        v[11] = t1;
        if (step < 5) {
            pa = na; pb = nb; pc = nc; pd = nd;
            const double det = pa * pd - pb * pc;
            if (pa < 0.0 || det < 0.0) ok = false;
            rdet = 1.0 / det;
        }
        __syncwarp();
    }
    return ok;
}

template <int UNROLL>
__device__ __forceinline__ bool gj_newton(double (&v)[12], double* sbuf) {
    const int lane = threadIdx.x & 31;
    bool ok = true;
#pragma unroll UNROLL
    for (int step = 0; step < 6; ++step) {
        if ((lane >> 1) == step) {
            double2* dst = reinterpret_cast<double2*>(sbuf + 12 * (lane & 1));
#pragma unroll
            for (int r = 0; r < 12; r += 2) dst[r >> 1] = make_double2(v[r], v[r + 1]);
        }
        __syncwarp();
        const double2 pk = *reinterpret_cast<const double2*>(sbuf);
        const double2 pk1 = *reinterpret_cast<const double2*>(sbuf + 12);
        const double det = pk.x * pk1.y - pk1.x * pk.y;
        if (pk.x < 0.0 || det < 0.0) ok = false;
        const double rdet = newton_rcp(det);
        const double t0 = (pk1.y * v[0] - pk1.x * v[1]) * rdet;
        const double t1 = (pk.x * v[1] - pk.y * v[0]) * rdet;
#pragma unroll
        for (int r = 2; r < 12; r += 2) {
            const double2 a = *reinterpret_cast<const double2*>(sbuf + r);
            const double2 b = *reinterpret_cast<const double2*>(sbuf + 12 + r);
            v[r - 2] = fma(-b.x, t1, fma(-a.x, t0, v[r]));
            v[r - 1] = fma(-b.y, t1, fma(-a.y, t0, v[r + 1]));
        }
        v[10] = t0;
        v[11] = t1;
        __syncwarp();
    }
    return ok;
}

// lookahead, rotated loop, double-buffered publish (sbuf: 2 x 32 doubles)
template <int UNROLL>
__device__ __forceinline__ bool gj_la_rot(double (&v)[12], double* sbuf) {
    const int lane = threadIdx.x & 31;
    bool ok = true;
    {
        const int role = lane >> 1;
        if (role == 0) {
            double2* dst = reinterpret_cast<double2*>(sbuf + 12 * (lane & 1));
#pragma unroll
            for (int r = 0; r < 12; r += 2) dst[r >> 1] = make_double2(v[r], v[r + 1]);
        } else if (role == 1) {
            double2* dst = reinterpret_cast<double2*>(sbuf + 24 + 4 * (lane & 1));
            dst[0] = make_double2(v[0], v[1]);
            dst[1] = make_double2(v[2], v[3]);
        }
    }
    __syncwarp();
    double pa, pb, pc, pd, rdet;
    {
        const double2 q0 = *reinterpret_cast<const double2*>(sbuf), q1 = *reinterpret_cast<const double2*>(sbuf + 12);
        pa = q0.x; pc = q0.y; pb = q1.x; pd = q1.y;
        const double det = pa * pd - pb * pc;
        if (pa < 0.0 || det < 0.0) ok = false;
        rdet = newton_rcp(det);
    }
#pragma unroll UNROLL
    for (int step = 0; step < 6; ++step) {
        const double* cur = sbuf + 32 * (step & 1);
        double* nxt = sbuf + 32 * ((step & 1) ^ 1);
        const double2 a2 = *reinterpret_cast<const double2*>(cur + 2);
        const double2 b2 = *reinterpret_cast<const double2*>(cur + 14);
        const double2 n00 = *reinterpret_cast<const double2*>(cur + 24), n01 = *reinterpret_cast<const double2*>(cur + 26);
        const double2 n10 = *reinterpret_cast<const double2*>(cur + 28), n11 = *reinterpret_cast<const double2*>(cur + 30);
        const double u0 = (pd * n00.x - pb * n00.y) * rdet, u1 = (pa * n00.y - pc * n00.x) * rdet;
        const double w0 = (pd * n10.x - pb * n10.y) * rdet, w1 = (pa * n10.y - pc * n10.x) * rdet;
        const double na = fma(-b2.x, u1, fma(-a2.x, u0, n01.x));
        const double nc = fma(-b2.y, u1, fma(-a2.y, u0, n01.y));
        const double nb = fma(-b2.x, w1, fma(-a2.x, w0, n11.x));
        const double nd = fma(-b2.y, w1, fma(-a2.y, w0, n11.y));
        const double ndet = na * nd - nb * nc;
        const double nrdet = newton_rcp(ndet);
        const double t0 = (pd * v[0] - pb * v[1]) * rdet;
        const double t1 = (pa * v[1] - pc * v[0]) * rdet;
        v[0] = fma(-b2.x, t1, fma(-a2.x, t0, v[2]));
        v[1] = fma(-b2.y, t1, fma(-a2.y, t0, v[3]));
#pragma unroll
        for (int r = 4; r < 12; r += 2) {
            const double2 a = *reinterpret_cast<const double2*>(cur + r);
            const double2 b = *reinterpret_cast<const double2*>(cur + 12 + r);
            v[r - 2] = fma(-b.x, t1, fma(-a.x, t0, v[r]));
            v[r - 1] = fma(-b.y, t1, fma(-a.y, t0, v[r + 1]));
        }
        v[10] = t0;
        v[11] = t1;
        if (step < 5) {
            if (na < 0.0 || ndet < 0.0) ok = false;
            pa = na; pb = nb; pc = nc; pd = nd; rdet = nrdet;
            const int role = (lane >> 1) - (step + 1);
            if (role == 0) {
                double2* dst = reinterpret_cast<double2*>(nxt + 12 * (lane & 1));
#pragma unroll
                for (int r = 0; r < 12; r += 2) dst[r >> 1] = make_double2(v[r], v[r + 1]);
            } else if (role == 1) {
                double2* dst = reinterpret_cast<double2*>(nxt + 24 + 4 * (lane & 1));
                dst[0] = make_double2(v[0], v[1]);
                dst[1] = make_double2(v[2], v[3]);
            }
        }
        __syncwarp();
    }
    return ok;
}


template <int VAR>
__global__ void __launch_bounds__(128, 5) k_gj(double* out, long long* cyc, const double* in, int nw, int reps) {
    __shared__ __align__(16) double sbuf[4 * 64];
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
            for (int r = 0; r < 12; ++r) col[r] = Q[r * 24 + (lane % 24)] + ((r == lane % 12 && lane < 12) ? 30.0 : 0.0) + acc * 1e-30;
            __syncwarp();
            bool ok;
            if (VAR == 0) ok = gauss_jordan12(col, sbuf + 64 * warp);
            else if (VAR == 1) ok = gj_newton<1>(col, sbuf + 64 * warp);
            else if (VAR == 2) ok = gauss_jordan12_la(col, sbuf + 64 * warp);
            else if (VAR == 3) ok = gj_la_rot<1>(col, sbuf + 64 * warp);
            else if (VAR == 4) ok = gj_la_rot<6>(col, sbuf + 64 * warp);
            else if (VAR == 5) ok = gj_newton<6>(col, sbuf + 64 * warp);
            else ok = gauss_jordan12_b4(col, sbuf + 64 * warp);
            if (it + 1 < reps) {
#pragma unroll
                for (int r = 0; r < 12; ++r) acc += col[r];
                acc += ok;
            }
        }
        long long t1 = clock64();
        if (lane == 0) cyc[blockIdx.x * 4 + warp] = (t1 - t0) / reps;
    }
    // result of the last elimination, for the cross-check
    for (int r = 0; r < 12; ++r) out[(size_t)blockIdx.x * 128 * 12 + threadIdx.x * 12 + r] = (warp < nw) ? col[r] : 0.0;
}

template <int VAR>
void run(double* out, long long* cyc, const double* in, double* ref) {
    for (int cfg = 0; cfg < 3; ++cfg) {
        const int bps = cfg == 0 ? 1 : 5, nw = cfg == 0 ? 1 : cfg == 1 ? 2 : 4;
        for (int rep = 0; rep < 2; ++rep) { k_gj<VAR><<<148 * bps, 128>>>(out, cyc, in, nw, 64); cudaDeviceSynchronize(); }
        long long s = 0; int n = 0;
        for (int b = 0; b < 148 * bps; ++b) for (int w = 0; w < nw; ++w) { s += cyc[b * 4 + w]; ++n; }
        double md = 0;
        if (VAR == 0 && cfg == 0) for (int i = 0; i < 32 * 12; ++i) ref[i] = out[i];
        for (int i = 0; i < 32 * 12; ++i) md = fmax(md, fabs(out[i] - ref[i]));
        printf("variant %d  blocks/SM %d  GJ warps/block %d : %5lld cycles per elimination (+tableau load)   max |x - x_v0| = %.3e\n", VAR, bps, nw, s / n, md);
    }
}
int main() {
    double *in, *out, *ref; long long* cyc;
    cudaMallocManaged(&in, 576 * 8); cudaMallocManaged(&out, (size_t)148 * 5 * 128 * 12 * 8); cudaMallocManaged(&cyc, 148 * 5 * 4 * 8);
    ref = (double*)malloc(32 * 12 * 8);
    for (int i = 0; i < 576; ++i) { const int r = i / 24, c = i % 24; in[i] = 0.01 * (((r < c ? r * 24 + c : c * 24 + r) * 7) % 13) + ((r == c) ? 1.0 : 0.0); }
    run<0>(out, cyc, in, ref); run<1>(out, cyc, in, ref); run<2>(out, cyc, in, ref); run<3>(out, cyc, in, ref); run<4>(out, cyc, in, ref); run<5>(out, cyc, in, ref); run<6>(out, cyc, in, ref);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
