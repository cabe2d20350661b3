// Backward sweep (Riccati recursion) and linear rollout of one problem, executed by
// one thread block with the working set in shared memory.
//
//   SinglePhase::backward_sweep          HSDDPSolver/source/SinglePhase.cpp:299-367
//   MultiPhaseDDP::backward_sweep        HSDDPSolver/source/MultiPhaseDDP.cpp:190-229
//   MultiPhaseDDP::impact_aware_step     :480-484
//   SinglePhase::linear_rollout          SinglePhase.cpp:145-178
//   MultiPhaseDDP::linear_rollout        MultiPhaseDDP.cpp:20-50
//
// Algebra of one stage (n = m = 24), with A = I + At (At non-zero in rows
// {0,1,2,6,7,8} and the three dt entries (3+j, 9+j)) and B = rows {6,7,8} dense over
// the GRF columns + scaled unit entries:
//     Y = H A,  Z = H B                          (sparse right factors: <= 7 FMA per entry)
//     Qxx = lxx + A^T Y, Qux = B^T Y, Quu = luu + B^T Z, Qx = lx + A^T Gn, Qu = lu + B^T Gn
//     Quu = L L^T (Cholesky);  PD test on Quu - 1e-9 I as the reference does (Q7)
//     W = L^-1 [Qux | Qu];  [K | dU] = -L^-T W;  H' = sym(Qxx) - W^T W;  G' = Qx - W^T w_u
// Skipping structurally zero terms of A and B leaves every retained product and its
// summation order unchanged, so this equals the dense arithmetic up to the
// association order of the sums.
#pragma once
#include "hsddp_device.cuh"

namespace hsddp {

struct PhaseConst {
    double cm[4];    // (c_l / m) dt   : B(9+j, 3l+j)
    double swdt[4];  // (1 - c_l) dt   : B(12+3l+j, 12+3l+j)
};

__device__ __forceinline__ PhaseConst phase_const(unsigned cmask, double dt) {
    PhaseConst pc;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const double c = (double)((cmask >> l) & 1u);
        pc.cm[l] = (c / hkd::kMass) * dt;
        pc.swdt[l] = (1.0 - c) * dt;
    }
    return pc;
}

__device__ __forceinline__ int rowsel(int r) { return r < 3 ? r : r + 3; }  // {0,1,2,6,7,8}

// (M A)[i,j] given column-major M
__device__ __forceinline__ double right_mul_A(const double* M, const double* At, double dt, int i, int j) {
    double v = M[i + 24 * j];
#pragma unroll
    for (int r = 0; r < 6; ++r) v += M[i + 24 * rowsel(r)] * At[r * 24 + j];
    if (j >= 9 && j < 12) v += M[i + 24 * (j - 6)] * dt;
    return v;
}
// (M B)[i,j]
__device__ __forceinline__ double right_mul_B(const double* M, const double* Bt, const PhaseConst& pc, int i, int j) {
    if (j < 12) {
        double v = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) v += M[i + 24 * (6 + a)] * Bt[a * 12 + j];
        v += M[i + 24 * (9 + j % 3)] * pc.cm[j / 3];
        return v;
    }
    return M[i + 24 * j] * pc.swdt[(j - 12) / 3];
}
// (A^T M)[i,j]
__device__ __forceinline__ double left_mul_At(const double* M, const double* At, double dt, int i, int j) {
    double v = M[i + 24 * j];
#pragma unroll
    for (int r = 0; r < 6; ++r) v += At[r * 24 + i] * M[rowsel(r) + 24 * j];
    if (i >= 9 && i < 12) v += dt * M[(i - 6) + 24 * j];
    return v;
}
// (B^T M)[i,j]
__device__ __forceinline__ double left_mul_Bt(const double* M, const double* Bt, const PhaseConst& pc, int i, int j) {
    if (i < 12) {
        double v = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) v += Bt[a * 12 + i] * M[(6 + a) + 24 * j];
        v += pc.cm[i / 3] * M[(9 + i % 3) + 24 * j];
        return v;
    }
    return pc.swdt[(i - 12) / 3] * M[i + 24 * j];
}
// vector versions
__device__ __forceinline__ double At_vec(const double* v, const double* At, double dt, int i) {
    double r = v[i];
#pragma unroll
    for (int q = 0; q < 6; ++q) r += At[q * 24 + i] * v[rowsel(q)];
    if (i >= 9 && i < 12) r += dt * v[i - 6];
    return r;
}
__device__ __forceinline__ double Bt_vec(const double* v, const double* Bt, const PhaseConst& pc, int i) {
    if (i < 12) {
        double r = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) r += Bt[a * 12 + i] * v[6 + a];
        r += pc.cm[i / 3] * v[9 + i % 3];
        return r;
    }
    return pc.swdt[(i - 12) / 3] * v[i];
}

// Cholesky of a 24x24 SPD matrix by one warp, lane i holding row i in registers.
// `shift` is subtracted from the diagonal first.  Returns false on a negative pivot
// (for shift != 0 this is the reference's LDLT(Quu - 1e-9 I).isPositive() verdict,
// by Sylvester's law of inertia).  When Lout != nullptr the factor is written
// column-major (lower triangle incl. diagonal).
__device__ inline bool warp_cholesky24(const double* A, double shift, double* Lout) {
    const int lane = threadIdx.x & 31;
    double a[24];
#pragma unroll
    for (int j = 0; j < 24; ++j) a[j] = (lane < 24) ? A[lane + 24 * j] : ((lane == j) ? 1.0 : 0.0);
    if (lane < 24) {
#pragma unroll
        for (int j = 0; j < 24; ++j) if (j == lane) a[j] -= shift;
    }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 24; ++k) {
        const double p = __shfl_sync(0xffffffffu, a[k], k);
        if (p < 0.0 || (shift == 0.0 && !(p > 0.0))) ok = false;
        const double sp = sqrt(fabs(p));
        const double r = (sp > 0.0) ? 1.0 / sp : 0.0;
        const double l = (lane == k) ? sp : a[k] * r;
        a[k] = l;
#pragma unroll
        for (int j = k + 1; j < 24; ++j) {
            const double lj = __shfl_sync(0xffffffffu, l, j);
            a[j] -= l * lj;
        }
    }
    if (Lout && lane < 24) {
#pragma unroll
        for (int j = 0; j < 24; ++j) if (j <= lane) Lout[lane + 24 * j] = a[j];
    }
    return ok;
}

// One phase of the backward sweep.  On entry sm.G / sm.H hold Gprime / Hprime (zero for
// the last phase).  Returns false if a stage failed the PD test.
__device__ inline bool phase_backward_sweep_block(Smem& sm, int ph, double reg, double& dV1, double& dV2) {
    const DevSchedule& sc = sm.sc;
    const int tid = threadIdx.x, warp = tid >> 5;
    const unsigned cm = sc.cmask[ph];
    const double dt = sc.dt;
    const PhaseConst pc = phase_const(cm, dt);
    const int Nph = sc.horizon[ph];
    const double* trec = sm.tq + ph * TQ_STRIDE;
    // G[N] = Phix + Gprime ; H[N] = Phixx + Hprime
    if (tid < 24) sm.G[tid] += trec[TQ_PHIX + tid];
    for (int e = tid; e < 576; e += kThreads) {
        const int i = e % 24, j = e / 24;
        double v = lxx_entry(i, j, cm, 0.0, 20.0, true);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const double wh = trec[TQ_WH + l];
            if (wh != 0.0) v += wh * (trec[TQ_HX + 24 * l + i] * trec[TQ_HX + 24 * l + j]);
        }
        sm.H[e] += v;
    }
    __syncthreads();
    dV1 = 0.0; dV2 = 0.0;
    for (int k = Nph - 1; k >= 0; --k) {
        const int s = sc.stage_off[ph] + k;
        const int n1 = sc.node_off[ph] + k + 1;
        // stage inputs
        for (int e = tid; e < LQ_STRIDE; e += kThreads) sm.lq[e] = sm.lqg[(size_t)s * LQ_STRIDE + e];
        if (tid < 24) sm.dfc[tid] = sm.Defect[24 * n1 + tid];
        __syncthreads();
        const double* At = sm.lq + LQ_AT;
        const double* Bt = sm.lq + LQ_BT;
        // Gn = G + H d   (Q10)
        if (tid < 24) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 24; ++j) acc += sm.H[tid + 24 * j] * sm.dfc[j];
            sm.Gn[tid] = sm.G[tid] + acc;
        }
        // Y = H A ; Z = H B
        for (int e = tid; e < 576; e += kThreads) {
            const int i = e % 24, j = e / 24;
            sm.Y[e] = right_mul_A(sm.H, At, dt, i, j);
            sm.Z[e] = right_mul_B(sm.H, Bt, pc, i, j);
        }
        __syncthreads();
        // Q function
        for (int e = tid; e < 576; e += kThreads) {
            const int i = e % 24, j = e / 24;
            double qxx = lxx_entry(i, j, cm, dt, dt, false) + left_mul_At(sm.Y, At, dt, i, j);
            double quu = left_mul_Bt(sm.Z, Bt, pc, i, j);
            if (i == j) { quu += dt * weight_R(i); qxx += reg; quu += reg; }
            if (i < 12 && j < 12 && i / 3 == j / 3) quu += sm.lq[LQ_LUU + 9 * (i / 3) + 3 * (i % 3) + (j % 3)];
            sm.Qxx[e] = qxx;
            sm.Quu[e] = quu;
            sm.Qux[e] = left_mul_Bt(sm.Y, Bt, pc, i, j);
        }
        if (tid < 24) {
            sm.Qx[tid] = sm.lq[LQ_LX + tid] + (At_vec(sm.Gn, At, dt, tid) - 0.0);
            sm.Qu[tid] = sm.lq[LQ_LU + tid] + Bt_vec(sm.Gn, Bt, pc, tid);
        }
        __syncthreads();
        // factorisation: warp 0 -> L (into Z, which is free now), warp 1 -> PD verdict of the shifted matrix
        if (warp == 0) {
            const bool ok = warp_cholesky24(sm.Quu, 0.0, sm.Z);
            if ((tid & 31) == 0) sm.ibuf[0] = ok ? 1 : 0;
        } else if (warp == 1) {
            const bool ok = warp_cholesky24(sm.Quu, 1e-9, nullptr);
            if ((tid & 31) == 0) sm.ibuf[1] = ok ? 1 : 0;
        }
        __syncthreads();
        if (!(sm.ibuf[0] && sm.ibuf[1])) return false;
        const double* L = sm.Z;
        // W = L^-1 [Qux | Qu] : one thread per right-hand side, result in place (Qux, wu)
        if (tid < 25) {
            double w[24];
            double* col = (tid < 24) ? (sm.Qux + 24 * tid) : sm.wu;
            const double* src = (tid < 24) ? col : sm.Qu;
#pragma unroll
            for (int i = 0; i < 24; ++i) {
                double sacc = src[i];
#pragma unroll
                for (int m = 0; m < i; ++m) sacc -= L[i + 24 * m] * w[m];
                w[i] = sacc / L[i + 24 * i];
            }
#pragma unroll
            for (int i = 0; i < 24; ++i) col[i] = w[i];
        }
        __syncthreads();
        const double* W = sm.Qux;
        // [K | dU] = -L^-T W  (threads 0..24)   ||   H' = sym(Qxx) - W^T W, G' = Qx - W^T wu (threads 32..127)
        if (tid < 25) {
            double xk[24];
            const double* col = (tid < 24) ? (W + 24 * tid) : sm.wu;
#pragma unroll
            for (int i = 23; i >= 0; --i) {
                double sacc = col[i];
#pragma unroll
                for (int m = i + 1; m < 24; ++m) sacc -= L[m + 24 * i] * xk[m];
                xk[i] = sacc / L[i + 24 * i];
            }
            if (tid < 24) {
                double* Kk = sm.K + 576 * (size_t)s + 24 * tid;
#pragma unroll
                for (int i = 0; i < 24; ++i) Kk[i] = -xk[i];
            } else {
                double dvk = 0.0;
#pragma unroll
                for (int i = 0; i < 24; ++i) { sm.dU[24 * s + i] = -xk[i]; dvk += sm.Qu[i] * xk[i]; }  // dV_k = -Qu^T dU
                sm.dbuf[0] = dvk;
            }
        } else if (tid >= 32) {
            for (int e = tid - 32; e < 576 + 24; e += kThreads - 32) {
                if (e < 576) {
                    const int i = e % 24, j = e / 24;
                    double acc = 0.0;
#pragma unroll
                    for (int m = 0; m < 24; ++m) acc += W[m + 24 * i] * W[m + 24 * j];
                    sm.H[e] = (sm.Qxx[i + 24 * j] + sm.Qxx[j + 24 * i]) / 2 - acc;
                } else {
                    const int i = e - 576;
                    double acc = 0.0;
#pragma unroll
                    for (int m = 0; m < 24; ++m) acc += W[m + 24 * i] * sm.wu[m];
                    sm.G[i] = sm.Qx[i] - acc;
                }
            }
        }
        __syncthreads();
        const double dvk = sm.dbuf[0];
        dV1 -= dvk;
        dV2 += dvk;
    }
    // G[0] += H[0] * Defect[0]
    {
        const int n0 = sc.node_off[ph];
        if (tid < 24) sm.dfc[tid] = sm.Defect[24 * n0 + tid];
        __syncthreads();
        double acc = 0.0;
        if (tid < 24) {
#pragma unroll
            for (int j = 0; j < 24; ++j) acc += sm.H[tid + 24 * j] * sm.dfc[j];
        }
        __syncthreads();
        if (tid < 24) sm.G[tid] += acc;
        __syncthreads();
    }
    return true;
}

// MultiPhaseDDP::backward_sweep(regularization)
__device__ inline bool backward_sweep_block(Smem& sm, double reg) {
    const DevSchedule& sc = sm.sc;
    const int tid = threadIdx.x;
    double dV1 = 0.0, dV2 = 0.0;
    bool success = true;
    for (int ph = sc.n_phases - 1; ph >= 0; --ph) {
        if (ph == sc.n_phases - 1) {
            for (int e = tid; e < 576; e += kThreads) sm.H[e] = 0.0;
            if (tid < 24) sm.G[tid] = 0.0;
            __syncthreads();
        } else {
            // impact-aware step: G' = Px^T G0, H' = Px^T H0 Px at the phase's terminal state
            const int ne = sc.node_off[ph] + sc.horizon[ph];
            resetmap_partial_block(sm.X + 24 * ne, sc.cmask[ph], sc.nmask[ph], sm.Y);
            const double* P = sm.Y;
            for (int e = tid; e < 576; e += kThreads) {  // Z = P^T H
                const int i = e % 24, j = e / 24;
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < 24; ++m) acc += P[m + 24 * i] * sm.H[m + 24 * j];
                sm.Z[e] = acc;
            }
            if (tid < 24) {
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < 24; ++m) acc += P[m + 24 * tid] * sm.G[m];
                sm.vtmp[tid] = acc;
            }
            __syncthreads();
            for (int e = tid; e < 576; e += kThreads) {  // H = Z P
                const int i = e % 24, j = e / 24;
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < 24; ++m) acc += sm.Z[i + 24 * m] * P[m + 24 * j];
                sm.H[e] = acc;
            }
            if (tid < 24) sm.G[tid] = sm.vtmp[tid];
            __syncthreads();
        }
        double d1, d2;
        if (!phase_backward_sweep_block(sm, ph, reg, d1, d2)) { success = false; break; }
        dV1 += d1;
        dV2 += d2;
    }
    if (success) {
        for (int e = tid; e < 576; e += kThreads) sm.g0h0[24 + e] = sm.H[e];
        if (tid < 24) sm.g0h0[tid] = sm.G[tid];
    }
    if (tid == 0) { sm.st.dV_1 = dV1; sm.st.dV_2 = dV2; sm.st.sweep_ok = success ? 1 : 0; }
    __syncthreads();
    return success;
}

// MultiPhaseDDP::backward_sweep_regularized (Q8).  Returns success; counts sweeps.
__device__ inline bool backward_sweep_regularized_block(Smem& sm, int& n_sweeps) {
    bool success = false;
    double reg = sm.st.reg;
    n_sweeps = 0;
    while (!success) {
        ++n_sweeps;
        success = backward_sweep_block(sm, reg);
        if (success) break;
        reg = fmax(reg * sm.opt.update_regularization, 1e-03);
        if (reg > 1e2) break;
    }
    reg = reg / 20;
    if (reg < 1e-06) reg = 0;
    if (threadIdx.x == 0) sm.st.reg = reg;
    __syncthreads();
    return success;
}

// ---------------------------------------------------------------------------
// linear rollout: dX recursion and expected cost change, warp 0 (lane i <-> component i)
// ---------------------------------------------------------------------------
__device__ inline void linear_rollout_block(Smem& sm, double eps) {
    const DevSchedule& sc = sm.sc;
    const int tid = threadIdx.x, lane = tid & 31;
    const double dt = sc.dt;
    double* sdx = sm.vtmp;   // current dx, shared for broadcast
    double* sdu = sm.vtmp2;  // current du
    double dV1 = 0.0, dV2 = 0.0;  // lane partial sums
    for (int ph = 0; ph < sc.n_phases; ++ph) {
        const unsigned cm = sc.cmask[ph];
        const PhaseConst pc = phase_const(cm, dt);
        const int n0 = sc.node_off[ph];
        // dx_init = Px dX_end(prev)
        if (ph > 0) {
            const int ne = sc.node_off[ph - 1] + sc.horizon[ph - 1];
            resetmap_partial_block(sm.X + 24 * ne, sc.cmask[ph - 1], sc.nmask[ph - 1], sm.Y);
            if (tid < 24) {
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < 24; ++m) acc += sm.Y[tid + 24 * m] * sdx[m];
                sm.Gn[tid] = acc;
            }
            __syncthreads();
        } else {
            if (tid < 24) sm.Gn[tid] = 0.0;
            __syncthreads();
        }
        if (tid < 32) {
            double dx = 0.0;
            if (lane < 24) {
                dx = sm.Gn[lane] + eps * sm.Defect[24 * n0 + lane];
                sm.dX[24 * n0 + lane] = dx;
                sdx[lane] = dx;
            }
            __syncwarp();
            for (int k = 0; k < sc.horizon[ph]; ++k) {
                const int s = sc.stage_off[ph] + k;
                const double* rec = sm.lqg + (size_t)s * LQ_STRIDE;
                const double* At = rec + LQ_AT;
                const double* Bt = rec + LQ_BT;
                double du = 0.0;
                if (lane < 24) {
                    const double* Kk = sm.K + 576 * (size_t)s;
                    double acc = 0.0;
#pragma unroll 8
                    for (int j = 0; j < 24; ++j) acc += Kk[lane + 24 * j] * sdx[j];
                    du = eps * sm.dU[24 * s + lane] + acc;
                    sdu[lane] = du;
                }
                __syncwarp();
                double dxn = 0.0;
                if (lane < 24) {
                    // A dx
                    double adx = dx;
                    if (lane < 3 || (lane >= 6 && lane < 9)) {
                        const int r = lane < 3 ? lane : lane - 3;
                        double acc = 0.0;
#pragma unroll 8
                        for (int j = 0; j < 24; ++j) acc += At[r * 24 + j] * sdx[j];
                        adx += acc;
                    } else if (lane >= 3 && lane < 6) {
                        adx += dt * sdx[lane + 6];
                    }
                    // B du
                    double bdu;
                    if (lane >= 6 && lane < 9) {
                        double acc = 0.0;
#pragma unroll
                        for (int j = 0; j < 12; ++j) acc += Bt[(lane - 6) * 12 + j] * sdu[j];
                        bdu = acc;
                    } else if (lane >= 9 && lane < 12) {
                        double acc = 0.0;
#pragma unroll
                        for (int l = 0; l < 4; ++l) acc += pc.cm[l] * sdu[3 * l + lane - 9];
                        bdu = acc;
                    } else if (lane >= 12) {
                        bdu = pc.swdt[(lane - 12) / 3] * du;
                    } else {
                        bdu = 0.0;
                    }
                    dxn = (adx + bdu) + eps * sm.Defect[24 * (n0 + k + 1) + lane];
                    // expected cost change, lane-partial
                    dV1 += rec[LQ_LX + lane] * dx + rec[LQ_LU + lane] * du;
                    // dx^T lxx dx
                    double qdx = 0.0;
                    {
                        qdx = lxx_entry(lane, lane, cm, dt, dt, false) * dx;
                        if (lane >= 3 && lane < 6) {
                            for (int l = 0; l < 4; ++l) qdx += lxx_entry(lane, 12 + 3 * l + lane - 3, cm, dt, dt, false) * sdx[12 + 3 * l + lane - 3];
                        } else if (lane >= 12) {
                            const int jj = (lane - 12) % 3;
                            qdx += lxx_entry(lane, 3 + jj, cm, dt, dt, false) * sdx[3 + jj];
                        }
                    }
                    dV2 += dx * qdx;
                    // du^T luu du
                    double rdu = (dt * weight_R(lane)) * du;
                    if (lane < 12) {
                        const int l = lane / 3, a = lane % 3;
#pragma unroll
                        for (int b = 0; b < 3; ++b) rdu += rec[LQ_LUU + 9 * l + 3 * a + b] * sdu[3 * l + b];
                    }
                    dV2 += du * rdu;
                }
                __syncwarp();
                dx = dxn;
                if (lane < 24) { sm.dX[24 * (n0 + k + 1) + lane] = dx; sdx[lane] = dx; }
                __syncwarp();
            }
            // terminal terms
            if (lane < 24) {
                const double* trec = sm.tq + ph * TQ_STRIDE;
                dV1 += trec[TQ_PHIX + lane] * dx;
                double qdx = lxx_entry(lane, lane, cm, 0.0, 20.0, true) * dx;
                if (lane >= 3 && lane < 6) {
                    for (int l = 0; l < 4; ++l) qdx += lxx_entry(lane, 12 + 3 * l + lane - 3, cm, 0.0, 20.0, true) * sdx[12 + 3 * l + lane - 3];
                } else if (lane >= 12) {
                    const int jj = (lane - 12) % 3;
                    qdx += lxx_entry(lane, 3 + jj, cm, 0.0, 20.0, true) * sdx[3 + jj];
                }
                for (int l = 0; l < 4; ++l) {
                    const double wh = trec[TQ_WH + l];
                    if (wh != 0.0) {
                        double hd = 0.0;
                        for (int j = 0; j < 24; ++j) hd += trec[TQ_HX + 24 * l + j] * sdx[j];
                        qdx += wh * trec[TQ_HX + 24 * l + lane] * hd;
                    }
                }
                dV2 += dx * qdx;
            }
        }
        __syncthreads();
    }
    if (tid < 32) {
        dV1 = warp_sum(dV1);
        dV2 = warp_sum(dV2);
        if (lane == 0) { sm.st.dV_1 = dV1; sm.st.dV_2 = dV2; }
    }
    __syncthreads();
}

}  // namespace hsddp
