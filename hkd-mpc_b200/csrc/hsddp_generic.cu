// Model-independent sweeps of SinglePhase<double, xs, us, ys> for the reference's OTHER instantiations
// (SURVEY.md 8f N4): <12,12,0> and <36,12,12> (HSDDPSolver/source/SinglePhase.cpp:538-540), and <24,24,0> in its
// dense form as a cross-check of the structure-exploiting HKD kernels.
//
// The reference ships a model, costs and a problem only for <24,24,0>; a device solver cannot call host plug-ins.
// The boundary of these instantiations is therefore the phase's storage after LQ_approximation
// (SinglePhase.cpp:265-296): A, B, C, D, RCostData {lx, lu, ly, lxx, luu, lux, lyy}, TCostData {Phix, Phixx} and
// Defect go in; backward_sweep (SinglePhase.cpp:299-367, incl. the ys > 0 output terms :329-336) and linear_rollout
// (SinglePhase.cpp:145-178) run on the device for a batch of independent phases; dU, K, G, H, dX, dV_1, dV_2 come out.
//
// One thread block per problem, everything of a stage in shared memory, FP64 throughout.  Dense products reuse
// Y = H A, Z = H B (the reference forms A^T H and B^T H and multiplies again: same values up to rounding); each
// thread owns a 3 x 4 register tile of an output (all sizes of the three instantiations are multiples of 12).
// Quu^-1 by Gauss-Jordan without pivoting on [Quu | I] (Quu is tested positive definite first).  PD verdict
// LDLT(Quu - 1e-9 I).isPositive() (:342-347) as in the HKD kernels: a non-positive pivot of Quu => false; all pivots
// positive and ||Quu^-1||_F < 5e8 => lambda_min(Quu) > 2e-9 => true; otherwise the shifted matrix is eliminated exactly.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../include/hsddp_b200.h"

namespace hsddp { void set_last_error(const std::string& s); }

namespace hsddp_generic {

struct PhasePtrs {
    int N, n;
    const double *A, *B, *C, *D, *lx, *lu, *ly, *lxx, *luu, *lux, *lyy, *Phix, *Phixx, *Defect;
    const double *Gprime, *Hprime, *dx_init;  // may be null (zeros)
    double *dU, *K, *G, *H, *dX, *dV;
    int* ok;
};

template <int XS, int US, int YS>
struct Lay {
    static constexpr int T = (XS >= 24) ? 128 : 32;  // threads per block
    static constexpr int XX = XS * XS, XU = XS * US, UU = US * US, YX = YS * XS, YU = YS * US, YY = YS * YS;
    // offsets in doubles; [A | B] and [Y | Z] are contiguous so that H [A | B] is one product
    static constexpr int oH = 0, oAB = oH + XX, oYZ = oAB + XX + XU, oQux = oYZ + XX + XU, oAug = oQux + XU, oQuu = oAug + 2 * UU,
                         oK = oQuu + UU, oC = oK + XU, oD = oC + YX, oLyy = oD + YU, oCtL = oLyy + YY, oDtL = oCtL + YX,
                         oVec = oDtL + YU, nVec = 5 * XS + 8 * US + YS + 40, total = oVec + nVec;
};

// Cout(i, j) = Cin(i, j) + sign * sum_l Aop(i, l) B(l, j), i < M, j < N, l < K (Cin null = 0);
// Aop(i, l) = TA ? A[l + lda i] : A[i + lda l]; B[l + ldb j].  3 x 4 register tiles, tile index strided over the block.
// The tile of Cin is read before the product and the tile of Cout written after it, so no memory access of the inner
// loop has to wait for a store that might alias it.
template <int T, int M, int N, int K, bool TA, bool NEG = false>
__device__ __forceinline__ void gemm_tiles(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                           const double* Cin, int ldi, double* Cout, int ldo, int tid) {
    constexpr int TM = 3, TN = 4, MT = M / TM, NT = N / TN;
    static_assert(M % TM == 0 && N % TN == 0, "sizes must be multiples of 12");
    for (int t = tid; t < MT * NT; t += T) {
        const int i0 = (t % MT) * TM, j0 = (t / MT) * TN;
        double acc[TM][TN], cin[TM][TN];
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c) { acc[r][c] = 0.0; cin[r][c] = Cin ? Cin[i0 + r + ldi * (j0 + c)] : 0.0; }
#pragma unroll 4
        for (int l = 0; l < K; ++l) {
            double a[TM], b[TN];
#pragma unroll
            for (int r = 0; r < TM; ++r) a[r] = TA ? A[l + lda * (i0 + r)] : A[i0 + r + lda * l];
#pragma unroll
            for (int c = 0; c < TN; ++c) b[c] = B[l + ldb * (j0 + c)];
#pragma unroll
            for (int r = 0; r < TM; ++r)
#pragma unroll
                for (int c = 0; c < TN; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
        }
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c) Cout[i0 + r + ldo * (j0 + c)] = NEG ? cin[r][c] - acc[r][c] : cin[r][c] + acc[r][c];
    }
}

template <int T>
__device__ __forceinline__ void copy_in(double* dst, const double* __restrict__ src, int n, int tid) {
    for (int e = tid; e < n; e += T) dst[e] = src[e];
}
// the same copy as 16-byte cp.async pieces (n even, both sides 16-byte aligned): fire and forget, one wait per stage
template <int T>
__device__ __forceinline__ void copy_in_async(double* dst, const double* __restrict__ src, int n, int tid) {
    for (int e = 2 * tid; e < n; e += 2 * T) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + e);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + e));
    }
}
__device__ __forceinline__ void copy_in_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// 1 / d for a positive pivot: hardware seed and two Newton steps (the IEEE division is ~80 instructions on the elimination's chain)
__device__ __forceinline__ double pivot_rcp(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
}

// SinglePhase::backward_sweep (SinglePhase.cpp:299-367)
template <int XS, int US, int YS>
__global__ void __launch_bounds__(Lay<XS, US, YS>::T) k_generic_backward_sweep(PhasePtrs p, double reg) {
    using L = Lay<XS, US, YS>;
    constexpr int T = L::T;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, pid = blockIdx.x, N = p.N;
    double* H = smem + L::oH;
    double* AB = smem + L::oAB;   // A (XS x XS) then B (XS x US), column-major
    double* YZ = smem + L::oYZ;   // H [A | B]
    double* Qux = smem + L::oQux; // US x XS
    double* Aug = smem + L::oAug; // US x 2 US: [Quu | I] -> [I | Quu^-1]
    double* Quu = smem + L::oQuu; // copy for the exact PD test
    double* Kk = smem + L::oK;    // US x XS
    double* Cm = smem + L::oC;
    double* Dm = smem + L::oD;
    double* Lyy = smem + L::oLyy;
    double* CtL = smem + L::oCtL; // XS x YS
    double* DtL = smem + L::oDtL; // US x YS
    double* G = smem + L::oVec;
    double* Gn = G + XS;
    double* dfc = Gn + XS;
    double* Qx = dfc + XS;
    double* Gnew = Qx + XS;
    double* Qu = Gnew + XS;
    double* dUk = Qu + US;
    double* colp = dUk + US;      // two buffers of US
    double* rowq = colp + 2 * US; // two buffers of 2 US
    double* ly = rowq + 4 * US;
    double* scal = ly + YS;  // [0] frob^2 partial / flags
    __shared__ int s_flag;

    const size_t pN = (size_t)pid * N, pN1 = (size_t)pid * (N + 1);
    // G[N] = Phix + G', H[N] = Phixx + H'
    for (int e = tid; e < L::XX; e += T) {
        const double v = p.Phixx[(size_t)pid * L::XX + e] + (p.Hprime ? p.Hprime[(size_t)pid * L::XX + e] : 0.0);
        H[e] = v;
        p.H[(pN1 + N) * L::XX + e] = v;
    }
    for (int e = tid; e < XS; e += T) {
        const double v = p.Phix[(size_t)pid * XS + e] + (p.Gprime ? p.Gprime[(size_t)pid * XS + e] : 0.0);
        G[e] = v;
        p.G[(pN1 + N) * XS + e] = v;
    }
    double dV1 = 0.0, dV2 = 0.0;
    bool success = true;
    __syncthreads();
    // stage inputs (A, B, C, D, lyy, ly, the defect of the next node) arrive by cp.async: the copies of stage k - 1 are issued
    // in the middle of stage k, as soon as its Q function is formed and the buffers are dead
    auto fetch_stage = [&](int k) {
        copy_in_async<T>(AB, p.A + (pN + k) * L::XX, L::XX, tid);
        copy_in_async<T>(AB + L::XX, p.B + (pN + k) * L::XU, L::XU, tid);
        copy_in_async<T>(dfc, p.Defect + (pN1 + k + 1) * XS, XS, tid);
        if (YS > 0) {
            copy_in_async<T>(Cm, p.C + (pN + k) * L::YX, L::YX, tid);
            copy_in_async<T>(Dm, p.D + (pN + k) * L::YU, L::YU, tid);
            copy_in_async<T>(Lyy, p.lyy + (pN + k) * L::YY, L::YY, tid);
            copy_in_async<T>(ly, p.ly + (pN + k) * YS, YS, tid);
        }
    };
    fetch_stage(N - 1);
    for (int k = N - 1; k >= 0; --k) {
        copy_in_wait();
        __syncthreads();
        // ---- [Y | Z] = H [A | B];  Gnext = G + H Defect[k+1];  C^T lyy, D^T lyy ----
        gemm_tiles<T, XS, XS + US, XS, false>(H, XS, AB, XS, nullptr, 0, YZ, XS, tid);
        for (int i = tid; i < XS; i += T) {
            double s = 0.0;
            for (int l = 0; l < XS; ++l) s = fma(H[i + XS * l], dfc[l], s);
            Gn[i] = G[i] + s;
        }
        if (YS > 0) {
            gemm_tiles<T, XS, YS, YS, true>(Cm, YS, Lyy, YS, nullptr, 0, CtL, XS, tid);
            gemm_tiles<T, US, YS, YS, true>(Dm, YS, Lyy, YS, nullptr, 0, DtL, US, tid);
        }
        __syncthreads();
        // ---- Q function.  The running-cost Hessians come straight from HBM as the products' initial values: every thread
        //      loads the 12 entries of its tile before its inner loop and needs them after it.  H is dead now: its storage
        //      takes Qxx ----
        for (int i = tid; i < XS + US; i += T) {
            double s = 0.0;
            if (i < XS) {
                for (int l = 0; l < XS; ++l) s = fma(AB[l + XS * i], Gn[l], s);
                s += p.lx[(pN + k) * XS + i];
                if (YS > 0) { double c = 0.0; for (int l = 0; l < YS; ++l) c = fma(Cm[l + YS * i], ly[l], c); s += c; }
                Qx[i] = s;
            } else {
                const int u = i - XS;
                for (int l = 0; l < XS; ++l) s = fma(AB[L::XX + l + XS * u], Gn[l], s);
                s += p.lu[(pN + k) * US + u];
                if (YS > 0) { double c = 0.0; for (int l = 0; l < YS; ++l) c = fma(Dm[l + YS * u], ly[l], c); s += c; }
                Qu[u] = s;
            }
        }
        gemm_tiles<T, XS, XS, XS, true>(AB, XS, YZ, XS, p.lxx + (pN + k) * L::XX, XS, H, XS, tid);
        gemm_tiles<T, US, XS, XS, true>(AB + L::XX, XS, YZ, XS, p.lux + (pN + k) * L::XU, US, Qux, US, tid);
        gemm_tiles<T, US, US, XS, true>(AB + L::XX, XS, YZ + L::XX, XS, p.luu + (pN + k) * L::UU, US, Aug, US, tid);
        if (YS > 0) {
            __syncthreads();
            gemm_tiles<T, XS, XS, YS, false>(CtL, XS, Cm, YS, H, XS, H, XS, tid);
            gemm_tiles<T, US, XS, YS, false>(DtL, US, Cm, YS, Qux, US, Qux, US, tid);
            gemm_tiles<T, US, US, YS, false>(DtL, US, Dm, YS, Aug, US, Aug, US, tid);
        }
        __syncthreads();
        if (k > 0) fetch_stage(k - 1);  // (A, B, C, D, lyy, ly, the defect are dead from here on)
        // regularisation; copy of Quu for the exact PD test; Qxx symmetrised in place (each thread owns the pairs (i, j), (j, i))
        for (int e = tid; e < L::XX; e += T) {
            const int i = e % XS, j = e / XS;
            if (i > j) continue;
            if (i == j) { H[e] += 1.0 * reg; continue; }
            const double v = (H[i + XS * j] + H[j + XS * i]) / 2;
            H[i + XS * j] = v;
            H[j + XS * i] = v;
        }
        for (int e = tid; e < L::UU; e += T) {
            const int i = e % US, j = e / US;
            double q = Aug[e];
            if (i == j) q += 1.0 * reg;
            Aug[e] = q;
            Quu[e] = q;
            Aug[L::UU + e] = (i == j) ? 1.0 : 0.0;
        }
        if (tid == 0) s_flag = 0;
        __syncthreads();
        // ---- Gauss-Jordan on [Quu | I]: thread (column j, row part h) keeps its RP entries of column j in registers;
        //      per step the owners publish column q and row q (two buffers alternate, so one barrier per step) ----
        {
            constexpr int PARTS = (T / (2 * US) >= 4) ? 4 : (T / (2 * US) >= 2) ? 2 : 1;
            constexpr int RP = US / PARTS;
            const int j = tid % (2 * US), h = tid / (2 * US);
            const bool active = tid < 2 * US * PARTS;
            double a[RP];
            if (active) {
#pragma unroll
                for (int i = 0; i < RP; ++i) a[i] = Aug[h * RP + i + US * j];
            }
            for (int q = 0; q < US; ++q) {
                double* cq = colp + (q & 1) * US;
                double* rq = rowq + (q & 1) * 2 * US;
                if (active) {
                    double mine = 0.0;
#pragma unroll
                    for (int i = 0; i < RP; ++i) {
                        if (j == q) cq[h * RP + i] = a[i];
                        if (h * RP + i == q) mine = a[i];
                    }
                    if (q / RP == h) rq[j] = mine;
                }
                __syncthreads();
                const double pv = cq[q];
                if (tid == 0 && !(pv > 0.0)) s_flag = 1;
                if (active) {
                    const double r = rq[j] * pivot_rcp(pv);
#pragma unroll
                    for (int i = 0; i < RP; ++i) a[i] = (h * RP + i == q) ? r : fma(-cq[h * RP + i], r, a[i]);
                }
            }
            __syncthreads();
            if (active) {
#pragma unroll
                for (int i = 0; i < RP; ++i) Aug[h * RP + i + US * j] = a[i];
            }
            __syncthreads();
        }
        // PD verdict of Quu - 1e-9 I
        if (s_flag == 0) {
            double f2 = 0.0;
            for (int e = tid; e < L::UU; e += T) { const double v = Aug[L::UU + e]; f2 = fma(v, v, f2); }
            for (int o = 16; o; o >>= 1) f2 += __shfl_xor_sync(0xffffffffu, f2, o);
            if ((tid & 31) == 0) scal[tid >> 5] = f2;
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int w = 0; w < T / 32; ++w) s += scal[w];
                if (!(s < 2.5e17)) s_flag = 2;  // ||Quu^-1||_F >= 5e8: decide exactly below
            }
            __syncthreads();
            if (s_flag == 2) {
                // exact: unpivoted LDL^T of Quu - 1e-9 I, positive iff every pivot is (rare path, one thread)
                if (tid == 0) {
                    int bad = 0;
                    for (int i = 0; i < US; ++i) Quu[i + US * i] -= 1.0 * 1e-9;
                    for (int q = 0; q < US && !bad; ++q) {
                        const double d = Quu[q + US * q];
                        if (!(d > 0.0)) { bad = 1; break; }
                        for (int i = q + 1; i < US; ++i) {
                            const double f = Quu[i + US * q] / d;
                            for (int jj = q + 1; jj <= i; ++jj) Quu[i + US * jj] -= f * Quu[jj + US * q];
                        }
                    }
                    s_flag = bad;
                }
                __syncthreads();
            }
        }
        if (s_flag != 0) { success = false; break; }
        // Quu_inv = (inv + inv^T) / 2 into the left block
        for (int e = tid; e < L::UU; e += T) {
            const int i = e % US, j = e / US;
            Aug[e] = (Aug[L::UU + i + US * j] + Aug[L::UU + j + US * i]) / 2;
        }
        __syncthreads();
        // ---- dU = -Quu_inv Qu; K = -Quu_inv Qux ----
        for (int i = tid; i < US; i += T) {
            double s = 0.0;
            for (int l = 0; l < US; ++l) s = fma(Aug[i + US * l], Qu[l], s);
            dUk[i] = -s;
        }
        gemm_tiles<T, US, XS, US, false, true>(Aug, US, Qux, US, nullptr, 0, Kk, US, tid);
        __syncthreads();
        // ---- G = Qx + Qux^T dU; H = sym(Qxx) + Qux^T K (into Y's storage, then copied back) ----
        gemm_tiles<T, XS, XS, US, true>(Qux, US, Kk, US, H, XS, YZ, XS, tid);
        for (int i = tid; i < XS; i += T) {
            double s = 0.0;
            for (int l = 0; l < US; ++l) s = fma(Qux[l + US * i], dUk[l], s);
            Gnew[i] = Qx[i] + s;
        }
        if (tid == 0) {
            double s = 0.0;
            for (int l = 0; l < US; ++l) s = fma(Qu[l], dUk[l], s);
            const double dV_k = -s;
            dV1 -= dV_k;
            dV2 += dV_k;
        }
        __syncthreads();
        for (int e = tid; e < L::XX; e += T) { const double v = YZ[e]; H[e] = v; p.H[(pN1 + k) * L::XX + e] = v; }
        for (int e = tid; e < XS; e += T) { const double v = Gnew[e]; G[e] = v; p.G[(pN1 + k) * XS + e] = v; }
        for (int e = tid; e < L::XU; e += T) p.K[(pN + k) * L::XU + e] = Kk[e];
        for (int e = tid; e < US; e += T) p.dU[(pN + k) * US + e] = dUk[e];
        __syncthreads();
    }
    // G[0] += H[0] Defect[0] — runs even after a failed stage, on whatever the storage holds (SinglePhase.cpp:365)
    copy_in_wait();  // (a failed stage leaves the copies of the next one in flight)
    __syncthreads();
    if (!success) {
        for (int e = tid; e < L::XX; e += T) H[e] = p.H[pN1 * L::XX + e];
        for (int e = tid; e < XS; e += T) G[e] = p.G[pN1 * XS + e];
    }
    copy_in<T>(dfc, p.Defect + pN1 * XS, XS, tid);
    __syncthreads();
    for (int i = tid; i < XS; i += T) {
        double s = 0.0;
        for (int l = 0; l < XS; ++l) s = fma(H[i + XS * l], dfc[l], s);
        p.G[pN1 * XS + i] = G[i] + s;
    }
    if (tid == 0) {
        p.dV[2 * pid] = dV1;
        p.dV[2 * pid + 1] = dV2;
        p.ok[pid] = success ? 1 : 0;
    }
}

// SinglePhase::linear_rollout (SinglePhase.cpp:145-178): one block of 64 threads per problem, thread i <-> component i
template <int XS, int US, int YS>
__global__ void __launch_bounds__(64) k_generic_linear_rollout(PhasePtrs p, double eps) {
    static_assert(XS <= 64 && US <= 64, "one thread per component");
    constexpr int XX = XS * XS, XU = XS * US, UU = US * US;
    __shared__ double dx[XS], du[US], red[4];
    const int tid = threadIdx.x, pid = blockIdx.x, N = p.N;
    const size_t pN = (size_t)pid * N, pN1 = (size_t)pid * (N + 1);
    if (tid < XS) {
        const double v = (p.dx_init ? p.dx_init[(size_t)pid * XS + tid] : 0.0) + eps * p.Defect[pN1 * XS + tid];
        dx[tid] = v;
        p.dX[pN1 * XS + tid] = v;
    }
    double dV1 = 0.0, dV2 = 0.0;
    __syncthreads();
    for (int k = 0; k < N; ++k) {
        const double* Kk = p.K + (pN + k) * XU;
        if (tid < US) {
            double s = 0.0;
            for (int j = 0; j < XS; ++j) s = fma(Kk[tid + US * j], dx[j], s);
            du[tid] = eps * p.dU[(pN + k) * US + tid] + s;
        }
        __syncthreads();
        double dxn = 0.0;
        if (tid < XS) {
            const double* Ak = p.A + (pN + k) * XX;
            const double* Bk = p.B + (pN + k) * XU;
            double a = 0.0, b = 0.0;
            for (int j = 0; j < XS; ++j) a = fma(Ak[tid + XS * j], dx[j], a);
            for (int j = 0; j < US; ++j) b = fma(Bk[tid + XS * j], du[j], b);
            dxn = (a + b) + eps * p.Defect[(pN1 + k + 1) * XS + tid];
            // expected cost change: lx dx, dx' lxx dx, (du' lux) dx
            const double* lxx = p.lxx + (pN + k) * XX;
            const double* lux = p.lux + (pN + k) * XU;
            double q = 0.0, pu = 0.0;
            for (int l = 0; l < XS; ++l) q = fma(lxx[l + XS * tid], dx[l], q);
            for (int l = 0; l < US; ++l) pu = fma(lux[l + US * tid], du[l], pu);
            dV1 = fma(p.lx[(pN + k) * XS + tid], dx[tid], dV1);
            dV2 += q * dx[tid] + pu * dx[tid];
        }
        if (tid < US) {
            const double* luu = p.luu + (pN + k) * UU;
            double r = 0.0;
            for (int l = 0; l < US; ++l) r = fma(luu[l + US * tid], du[l], r);
            dV1 = fma(p.lu[(pN + k) * US + tid], du[tid], dV1);
            dV2 += r * du[tid];
        }
        __syncthreads();
        if (tid < XS) { dx[tid] = dxn; p.dX[(pN1 + k + 1) * XS + tid] = dxn; }
        __syncthreads();
    }
    if (tid < XS) {
        const double* Phixx = p.Phixx + (size_t)pid * XX;
        double q = 0.0;
        for (int l = 0; l < XS; ++l) q = fma(Phixx[l + XS * tid], dx[l], q);
        dV1 = fma(p.Phix[(size_t)pid * XS + tid], dx[tid], dV1);
        dV2 += q * dx[tid];
    }
    for (int o = 16; o; o >>= 1) { dV1 += __shfl_xor_sync(0xffffffffu, dV1, o); dV2 += __shfl_xor_sync(0xffffffffu, dV2, o); }
    if ((tid & 31) == 0) { red[2 * (tid >> 5)] = dV1; red[2 * (tid >> 5) + 1] = dV2; }
    __syncthreads();
    if (tid == 0) { p.dV[2 * pid] = red[0] + red[2]; p.dV[2 * pid + 1] = red[1] + red[3]; }
}

}  // namespace hsddp_generic

// ===========================================================================
// C ABI
// ===========================================================================
using namespace hsddp_generic;

#define CKG(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            hsddp::set_last_error(std::string(#call) + ": " + cudaGetErrorString(e_));             \
            return HSDDP_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

struct hsddp_phase_batch {
    int device = 0, xs = 0, us = 0, ys = 0, N = 0, n = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double* in[HSDDP_PH_N_INPUTS] = {};
    double* out[HSDDP_PH_N_OUTPUTS] = {};
    double *Gprime = nullptr, *Hprime = nullptr, *dx_init = nullptr;
    int* ok = nullptr;
    bool have_gains = false;
};

static size_t ph_in_count(const hsddp_phase_batch* b, int which) {
    const size_t xs = b->xs, us = b->us, ys = b->ys, N = b->N, n = b->n;
    switch (which) {
        case HSDDP_PH_A: return n * N * xs * xs;
        case HSDDP_PH_B: return n * N * xs * us;
        case HSDDP_PH_C: return n * N * ys * xs;
        case HSDDP_PH_D: return n * N * ys * us;
        case HSDDP_PH_LX: return n * N * xs;
        case HSDDP_PH_LU: return n * N * us;
        case HSDDP_PH_LY: return n * N * ys;
        case HSDDP_PH_LXX: return n * N * xs * xs;
        case HSDDP_PH_LUU: return n * N * us * us;
        case HSDDP_PH_LUX: return n * N * us * xs;
        case HSDDP_PH_LYY: return n * N * ys * ys;
        case HSDDP_PH_PHIX: return n * xs;
        case HSDDP_PH_PHIXX: return n * xs * xs;
        case HSDDP_PH_DEFECT: return n * (N + 1) * xs;
    }
    return 0;
}
static size_t ph_out_count(const hsddp_phase_batch* b, int which) {
    const size_t xs = b->xs, us = b->us, N = b->N, n = b->n;
    switch (which) {
        case HSDDP_PH_OUT_DU: return n * N * us;
        case HSDDP_PH_OUT_K: return n * N * us * xs;
        case HSDDP_PH_OUT_G: return n * (N + 1) * xs;
        case HSDDP_PH_OUT_H: return n * (N + 1) * xs * xs;
        case HSDDP_PH_OUT_DX: return n * (N + 1) * xs;
        case HSDDP_PH_OUT_DV: return n * 2;
    }
    return 0;
}

template <int XS, int US, int YS>
static int launch_sweep(hsddp_phase_batch* b, const PhasePtrs& p, double reg) {
    using L = Lay<XS, US, YS>;
    const size_t bytes = (size_t)L::total * sizeof(double);
    CKG(cudaFuncSetAttribute(k_generic_backward_sweep<XS, US, YS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    CKG(cudaFuncSetAttribute(k_generic_backward_sweep<XS, US, YS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (getenv("HSDDP_VERBOSE")) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_generic_backward_sweep<XS, US, YS>, L::T, bytes);
        fprintf(stderr, "k_generic_backward_sweep<%d,%d,%d>: %d threads, %zu bytes of shared memory, %d blocks per SM\n", XS, US, YS, L::T, bytes, nb);
    }
    k_generic_backward_sweep<XS, US, YS><<<b->n, L::T, bytes, b->stream>>>(p, reg);
    return HSDDP_OK;
}
template <int XS, int US, int YS>
static int launch_rollout(hsddp_phase_batch* b, const PhasePtrs& p, double eps) {
    k_generic_linear_rollout<XS, US, YS><<<b->n, 64, 0, b->stream>>>(p, eps);
    return HSDDP_OK;
}
static PhasePtrs ph_ptrs(const hsddp_phase_batch* b) {
    PhasePtrs p{};
    p.N = b->N; p.n = b->n;
    p.A = b->in[HSDDP_PH_A]; p.B = b->in[HSDDP_PH_B]; p.C = b->in[HSDDP_PH_C]; p.D = b->in[HSDDP_PH_D];
    p.lx = b->in[HSDDP_PH_LX]; p.lu = b->in[HSDDP_PH_LU]; p.ly = b->in[HSDDP_PH_LY]; p.lxx = b->in[HSDDP_PH_LXX];
    p.luu = b->in[HSDDP_PH_LUU]; p.lux = b->in[HSDDP_PH_LUX]; p.lyy = b->in[HSDDP_PH_LYY]; p.Phix = b->in[HSDDP_PH_PHIX];
    p.Phixx = b->in[HSDDP_PH_PHIXX]; p.Defect = b->in[HSDDP_PH_DEFECT];
    p.dU = b->out[HSDDP_PH_OUT_DU]; p.K = b->out[HSDDP_PH_OUT_K]; p.G = b->out[HSDDP_PH_OUT_G]; p.H = b->out[HSDDP_PH_OUT_H];
    p.dX = b->out[HSDDP_PH_OUT_DX]; p.dV = b->out[HSDDP_PH_OUT_DV];
    p.ok = b->ok;
    return p;
}

extern "C" {

int hsddp_phase_batch_destroy(hsddp_phase_batch* b) {
    if (!b) return HSDDP_OK;
    cudaSetDevice(b->device);
    for (double* q : b->in) if (q) cudaFree(q);
    for (double* q : b->out) if (q) cudaFree(q);
    if (b->Gprime) cudaFree(b->Gprime);
    if (b->Hprime) cudaFree(b->Hprime);
    if (b->dx_init) cudaFree(b->dx_init);
    if (b->ok) cudaFree(b->ok);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
    return HSDDP_OK;
}

static int phase_batch_create_impl(int device, int xs, int us, int ys, int horizon, int n_problems, hsddp_phase_batch* b) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        hsddp::set_last_error("hsddp_phase_batch_create: no such CUDA device (there is no CPU fallback)");
        return HSDDP_ERR_CUDA;
    }
    CKG(cudaSetDevice(device));
    b->device = device; b->xs = xs; b->us = us; b->ys = ys; b->N = horizon; b->n = n_problems;
    CKG(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    CKG(cudaEventCreate(&b->ev0));
    CKG(cudaEventCreate(&b->ev1));
    for (int w = 0; w < HSDDP_PH_N_INPUTS; ++w) {
        const size_t c = ph_in_count(b, w);
        if (c == 0) continue;
        CKG(cudaMalloc(&b->in[w], c * sizeof(double)));
        CKG(cudaMemsetAsync(b->in[w], 0, c * sizeof(double), b->stream));
    }
    for (int w = 0; w < HSDDP_PH_N_OUTPUTS; ++w) {
        const size_t c = ph_out_count(b, w);
        CKG(cudaMalloc(&b->out[w], c * sizeof(double)));
        CKG(cudaMemsetAsync(b->out[w], 0, c * sizeof(double), b->stream));
    }
    CKG(cudaMalloc(&b->Gprime, (size_t)n_problems * xs * sizeof(double)));
    CKG(cudaMalloc(&b->Hprime, (size_t)n_problems * xs * xs * sizeof(double)));
    CKG(cudaMalloc(&b->dx_init, (size_t)n_problems * xs * sizeof(double)));
    CKG(cudaMalloc(&b->ok, (size_t)n_problems * sizeof(int)));
    CKG(cudaMemsetAsync(b->ok, 0, (size_t)n_problems * sizeof(int), b->stream));
    CKG(cudaStreamSynchronize(b->stream));
    return HSDDP_OK;
}

int hsddp_phase_batch_create(int device, int xs, int us, int ys, int horizon, int n_problems, hsddp_phase_batch** out) {
    if (!out) return HSDDP_ERR_ARG;
    *out = nullptr;
    const bool known = (xs == 24 && us == 24 && ys == 0) || (xs == 12 && us == 12 && ys == 0) || (xs == 36 && us == 12 && ys == 12);
    if (!known) {
        hsddp::set_last_error("hsddp_phase_batch_create: only the reference's instantiations <24,24,0>, <12,12,0>, <36,12,12> are built (SinglePhase.cpp:538-540)");
        return HSDDP_ERR_UNSUPPORTED;
    }
    if (horizon < 1 || n_problems < 1) { hsddp::set_last_error("hsddp_phase_batch_create: horizon and n_problems must be positive"); return HSDDP_ERR_ARG; }
    hsddp_phase_batch* b = nullptr;
    try { b = new hsddp_phase_batch(); } catch (...) { return HSDDP_ERR_ARG; }
    const int rc = phase_batch_create_impl(device, xs, us, ys, horizon, n_problems, b);
    if (rc != HSDDP_OK) { hsddp_phase_batch_destroy(b); return rc; }
    *out = b;
    return HSDDP_OK;
}

int hsddp_phase_batch_set(hsddp_phase_batch* b, int which, const double* host) {
    if (!b || !host || which < 0 || which >= HSDDP_PH_N_INPUTS) return HSDDP_ERR_ARG;
    const size_t c = ph_in_count(b, which);
    if (c == 0) return HSDDP_OK;  // C, D, ly, lyy of a ys == 0 instantiation
    CKG(cudaSetDevice(b->device));
    CKG(cudaMemcpyAsync(b->in[which], host, c * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    CKG(cudaStreamSynchronize(b->stream));
    return HSDDP_OK;
}

int hsddp_phase_batch_backward_sweep(hsddp_phase_batch* b, double regularization, const double* Gprime, const double* Hprime, int32_t* ok) {
    if (!b || !(regularization >= 0.0)) return HSDDP_ERR_ARG;
    CKG(cudaSetDevice(b->device));
    PhasePtrs p = ph_ptrs(b);
    if (Gprime) { CKG(cudaMemcpyAsync(b->Gprime, Gprime, (size_t)b->n * b->xs * sizeof(double), cudaMemcpyHostToDevice, b->stream)); p.Gprime = b->Gprime; }
    if (Hprime) { CKG(cudaMemcpyAsync(b->Hprime, Hprime, (size_t)b->n * b->xs * b->xs * sizeof(double), cudaMemcpyHostToDevice, b->stream)); p.Hprime = b->Hprime; }
    CKG(cudaEventRecord(b->ev0, b->stream));
    int rc;
    if (b->xs == 24) rc = launch_sweep<24, 24, 0>(b, p, regularization);
    else if (b->xs == 12) rc = launch_sweep<12, 12, 0>(b, p, regularization);
    else rc = launch_sweep<36, 12, 12>(b, p, regularization);
    if (rc != HSDDP_OK) return rc;
    CKG(cudaGetLastError());
    CKG(cudaEventRecord(b->ev1, b->stream));
    if (ok) CKG(cudaMemcpyAsync(ok, b->ok, (size_t)b->n * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CKG(cudaStreamSynchronize(b->stream));
    b->have_gains = true;
    return HSDDP_OK;
}

int hsddp_phase_batch_linear_rollout(hsddp_phase_batch* b, double eps, const double* dx_init) {
    if (!b) return HSDDP_ERR_ARG;
    if (!b->have_gains) { hsddp::set_last_error("hsddp_phase_batch_linear_rollout: no gains yet (run the backward sweep or set K, dU)"); return HSDDP_ERR_STATE; }
    CKG(cudaSetDevice(b->device));
    PhasePtrs p = ph_ptrs(b);
    if (dx_init) { CKG(cudaMemcpyAsync(b->dx_init, dx_init, (size_t)b->n * b->xs * sizeof(double), cudaMemcpyHostToDevice, b->stream)); p.dx_init = b->dx_init; }
    CKG(cudaEventRecord(b->ev0, b->stream));
    int rc;
    if (b->xs == 24) rc = launch_rollout<24, 24, 0>(b, p, eps);
    else if (b->xs == 12) rc = launch_rollout<12, 12, 0>(b, p, eps);
    else rc = launch_rollout<36, 12, 12>(b, p, eps);
    if (rc != HSDDP_OK) return rc;
    CKG(cudaGetLastError());
    CKG(cudaEventRecord(b->ev1, b->stream));
    CKG(cudaStreamSynchronize(b->stream));
    return HSDDP_OK;
}

int hsddp_phase_batch_get(hsddp_phase_batch* b, int which, double* host) {
    if (!b || !host || which < 0 || which >= HSDDP_PH_N_OUTPUTS) return HSDDP_ERR_ARG;
    CKG(cudaSetDevice(b->device));
    CKG(cudaMemcpyAsync(host, b->out[which], ph_out_count(b, which) * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    CKG(cudaStreamSynchronize(b->stream));
    return HSDDP_OK;
}

int hsddp_phase_batch_set_gains(hsddp_phase_batch* b, const double* dU, const double* K) {
    if (!b || !dU || !K) return HSDDP_ERR_ARG;
    CKG(cudaSetDevice(b->device));
    CKG(cudaMemcpyAsync(b->out[HSDDP_PH_OUT_DU], dU, ph_out_count(b, HSDDP_PH_OUT_DU) * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    CKG(cudaMemcpyAsync(b->out[HSDDP_PH_OUT_K], K, ph_out_count(b, HSDDP_PH_OUT_K) * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    CKG(cudaStreamSynchronize(b->stream));
    b->have_gains = true;
    return HSDDP_OK;
}

int hsddp_phase_batch_last_ms(hsddp_phase_batch* b, float* ms) {
    if (!b || !ms) return HSDDP_ERR_ARG;
    CKG(cudaSetDevice(b->device));
    CKG(cudaEventElapsedTime(ms, b->ev0, b->ev1));
    return HSDDP_OK;
}

}  // extern "C"
