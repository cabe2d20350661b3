// Backward sweep (Riccati recursion) and linear rollout of one problem, executed by
// one 4-warp thread block with the working set in shared memory and registers.
//
//   SinglePhase::backward_sweep          HSDDPSolver/source/SinglePhase.cpp:299-367
//   MultiPhaseDDP::backward_sweep        HSDDPSolver/source/MultiPhaseDDP.cpp:190-229
//   MultiPhaseDDP::impact_aware_step     :480-484
//   SinglePhase::linear_rollout          SinglePhase.cpp:145-178
//   MultiPhaseDDP::linear_rollout        MultiPhaseDDP.cpp:20-50
//
// Structure exploited (all of it exact: only structurally-zero terms are skipped):
//   * A = I + At with At non-zero in rows 0..11 only (rows 9..11 are zero, rows 3..5 hold dt)
//   * per leg exactly ONE 3-vector of controls is coupled to the state: the GRF of a
//     stance leg or the joint-velocity command of a swing leg.  The other 12 controls
//     see B = 0, so Quu is [Quu_r (12x12) ; diag(dt R + reg)] and K has 12 non-zero rows.
//     "Reduced" control index c = 3*leg + j  <->  full index 3*leg+j (stance) / 12+3*leg+j (swing).
//   * B_r (24x12): stance columns have rows {6,7,8} (torque arm) and row 9+j (1/m);
//     swing columns have the single entry (12+3l+j) = dt.
// One stage:
//     Y = H A            Z = H B_r                              (DMMA m8n8k4, K = 12 / 8)
//     Qxx = lxx + A^T Y  Qux_r = B_r^T Y   Quu_r = luu_r + B_r^T Z    Qx, Qu_r   (DMMA + fix-ups)
//     Gauss-Jordan on the register tableau [Quu_r | Qux_r | Qu_r] (lane = column) -> -K_r, -dU_r
//     PD verdict = no negative pivot of Quu_r - 1e-9 I (third warp, concurrently; Q7)
//     H' = sym(Qxx) + Qux_r^T K_r      G' = Qx + Qux_r^T dU_r   (DMMA, accumulators kept in registers)
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

// full control index of reduced index c
__device__ __forceinline__ int act_index(int c, unsigned cmask) { return ((cmask >> (c / 3)) & 1u) ? c : 12 + c; }
// full control index of the c-th inactive control
__device__ __forceinline__ int inact_index(int c, unsigned cmask) { return ((cmask >> (c / 3)) & 1u) ? 12 + c : c; }

// D(8x8) += A(8x4) * B(4x8), FP64 tensor-core tile.  Fragment layout (PTX ISA, m8n8k4 .f64):
//   a: row = lane>>2, col = lane&3 ; b: row = lane&3, col = lane>>2 ; c[i]: row = lane>>2, col = 2*(lane&3)+i
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// Shared-memory views used by the sweep (all with row stride 24 doubles):
//   H    [24][24]  symmetric value Hessian
//   Y    [24][24]  H A                         (row-major)
//   Zr   [24][24]  H B_r in columns 0..11 (12..15 zero padding)
//   At12 [12][24]  rows 0..11 of A - I
//   Bq   [ 8][24]  rows 4..11 of B_r in columns 0..11 (stance columns only; 12..15 zero)
//   QuxR [16][24]  Qux_r (rows 12..15 padding)
//   QuuR [16][24]  Quu_r in [0..11][0..11]
//   KrS  [12][24]  K_r of the current stage
struct SweepSmem {
    double *H, *Y, *Zr, *At12, *Bq, *QuxR, *QuuR, *KrS;
};

__device__ __forceinline__ SweepSmem sweep_views(Smem& sm) {
    SweepSmem v;
    v.H = sm.H; v.Y = sm.Y; v.Zr = sm.Z; v.QuxR = sm.Qux; v.QuuR = sm.Quu;
    v.At12 = sm.Qxx; v.Bq = sm.Qxx + 288; v.KrS = sm.Qxx + 288;  // KrS aliases Bq + tail: Bq is dead after the Q phase
    return v;
}

// Gauss-Jordan on 12 rows, one tableau column per lane (v[0..11]).  Lanes 0..11 of the
// warp must hold the columns of the 12x12 pivot matrix.  `sbuf` is a 12-double per-warp
// staging area.  Returns false if a negative pivot was met (only meaningful for the caller
// that runs the shifted matrix).  After the call v = Quu_r^-1 * (original column).
__device__ __forceinline__ bool gauss_jordan12(double (&v)[12], double* sbuf) {
    const int lane = threadIdx.x & 31;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        if (lane == k) {
            const double inv = 1.0 / v[k];
#pragma unroll
            for (int r = 0; r < 12; ++r) sbuf[r] = (r == k) ? inv : v[r] * inv;  // multipliers g_r = a_rk / a_kk
            sbuf[12] = v[k];
        }
        __syncwarp();
        double g[12];
#pragma unroll
        for (int r = 0; r < 12; r += 2) {
            const double2 t = *reinterpret_cast<const double2*>(sbuf + r);
            g[r] = t.x; g[r + 1] = t.y;
        }
        const double piv = sbuf[12];
        if (piv < 0.0) ok = false;
        const double vk = v[k];
#pragma unroll
        for (int r = 0; r < 12; ++r) v[r] = (r == k) ? vk * g[k] : fma(-g[r], vk, v[r]);
        __syncwarp();
    }
    return ok;
}

// One phase of the backward sweep.  On entry sm.G / sm.H hold Gprime / Hprime (zero for
// the last phase).  Returns false if a stage failed the PD test.
__device__ inline bool phase_backward_sweep_block(Smem& sm, int ph, double reg, double& dV1, double& dV2) {
    const DevSchedule& sc = sm.sc;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const unsigned cm = sc.cmask[ph];
    const double dt = sc.dt;
    const PhaseConst pc = phase_const(cm, dt);
    const int Nph = sc.horizon[ph];
    const double* trec = sm.tq + ph * TQ_STRIDE;
    const SweepSmem v = sweep_views(sm);
    // G[N] = Phix + Gprime ; H[N] = Phixx + Hprime
    if (tid < 24) sm.G[tid] += trec[TQ_PHIX + tid];
    for (int e = tid; e < 576; e += kThreads) {
        const int i = e / 24, j = e % 24;
        double val = lxx_entry(i, j, cm, 0.0, 20.0, true);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const double wh = trec[TQ_WH + l];
            if (wh != 0.0) val += wh * (trec[TQ_HX + 24 * l + i] * trec[TQ_HX + 24 * l + j]);
        }
        v.H[e] += val;
    }
    __syncthreads();
    dV1 = 0.0; dV2 = 0.0;
    for (int k = Nph - 1; k >= 0; --k) {
        const int s = sc.stage_off[ph] + k;
        const int n1 = sc.node_off[ph] + k + 1;
        const double* rec = sm.lqg + (size_t)s * LQ_STRIDE;
        // ---- stage inputs -> shared memory ----
        for (int e = tid; e < 288; e += kThreads) {  // At12: rows 0..11 of A - I
            const int r = e / 24, c = e % 24;
            double val = 0.0;
            if (r < 3) val = rec[LQ_AT + r * 24 + c];
            else if (r < 6) val = (c == r + 6) ? dt : 0.0;
            else if (r < 9) val = rec[LQ_AT + (r - 3) * 24 + c];
            v.At12[e] = val;
        }
        for (int e = tid; e < 8 * 16; e += kThreads) {  // Bq: rows 4..11 of B_r, reduced columns 0..15
            const int r = e / 16 + 4, c = e % 16;
            double val = 0.0;
            if (c < 12 && ((cm >> (c / 3)) & 1u)) {
                if (r >= 6 && r < 9) val = rec[LQ_BT + (r - 6) * 12 + c];
                else if (r == 9 + c % 3) val = pc.cm[c / 3];
            }
            v.Bq[(r - 4) * 24 + c] = val;
        }
        if (tid < 24) {
            sm.dfc[tid] = sm.Defect[24 * n1 + tid];
            sm.lq[LQ_LX + tid] = rec[LQ_LX + tid];
            sm.lq[LQ_LU + tid] = rec[LQ_LU + tid];
        } else if (tid >= 32 && tid < 68) {
            sm.lq[LQ_LUU + tid - 32] = rec[LQ_LUU + tid - 32];
        }
        __syncthreads();
        // ---- P1: Y = H A, Z = H B_r, Gn = G + H d ----
        if (warp < 3) {
            const int i0 = 8 * warp;
            double cy[3][2], cz[2] = {0.0, 0.0};
#pragma unroll
            for (int J = 0; J < 3; ++J) {
                const double2 h2 = *reinterpret_cast<const double2*>(v.H + (i0 + g) * 24 + 8 * J + 2 * t);
                cy[J][0] = h2.x; cy[J][1] = h2.y;
            }
#pragma unroll
            for (int kk = 0; kk < 12; kk += 4) {
                const double a = v.H[(kk + t) * 24 + i0 + g];  // H[i][k] by symmetry
#pragma unroll
                for (int J = 0; J < 3; ++J) dmma884(cy[J], a, v.At12[(kk + t) * 24 + 8 * J + g]);
                if (kk >= 4) dmma884(cz, a, v.Bq[(kk - 4 + t) * 24 + g]);
            }
#pragma unroll
            for (int J = 0; J < 3; ++J)
                *reinterpret_cast<double2*>(v.Y + (i0 + g) * 24 + 8 * J + 2 * t) = make_double2(cy[J][0], cy[J][1]);
            // swing columns of Z: H[:, 12+3l+j] * dt
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int c = 2 * t + q;
                if (!((cm >> (c / 3)) & 1u)) cz[q] += v.H[(i0 + g) * 24 + 12 + c] * pc.swdt[c / 3];
            }
            *reinterpret_cast<double2*>(v.Zr + (i0 + g) * 24 + 2 * t) = make_double2(cz[0], cz[1]);
        } else {
            // warp 3: Z columns 8..15 for all three row blocks, then Gn
#pragma unroll
            for (int I = 0; I < 3; ++I) {
                const int i0 = 8 * I;
                double cz[2] = {0.0, 0.0};
#pragma unroll
                for (int kk = 4; kk < 12; kk += 4) dmma884(cz, v.H[(kk + t) * 24 + i0 + g], v.Bq[(kk - 4 + t) * 24 + 8 + g]);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int c = 8 + 2 * t + q;
                    if (c < 12 && !((cm >> (c / 3)) & 1u)) cz[q] += v.H[(i0 + g) * 24 + 12 + c] * pc.swdt[c / 3];
                }
                *reinterpret_cast<double2*>(v.Zr + (i0 + g) * 24 + 8 + 2 * t) = make_double2(cz[0], cz[1]);
            }
            if (lane < 24) {  // Gn = G + H d   (Q10)
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int j = 0; j < 24; j += 4) {
                    a0 = fma(v.H[j * 24 + lane], sm.dfc[j], a0);
                    a1 = fma(v.H[(j + 1) * 24 + lane], sm.dfc[j + 1], a1);
                    a2 = fma(v.H[(j + 2) * 24 + lane], sm.dfc[j + 2], a2);
                    a3 = fma(v.H[(j + 3) * 24 + lane], sm.dfc[j + 3], a3);
                }
                sm.Gn[lane] = sm.G[lane] + ((a0 + a1) + (a2 + a3));
            }
        }
        __syncthreads();
        // ---- P2: Qxx (lower tiles, kept in registers), Qux_r, Quu_r, Qx, Qu_r ----
        // Qxx tile ownership: warp0 (0,0),(1,0) ; warp1 (1,1),(2,0) ; warp2 (2,1) ; warp3 (2,2)
        double cq[2][2];
        int qi[2], qj[2];
        int nq;
        if (warp == 0) { nq = 2; qi[0] = 0; qj[0] = 0; qi[1] = 1; qj[1] = 0; }
        else if (warp == 1) { nq = 2; qi[0] = 1; qj[0] = 1; qi[1] = 2; qj[1] = 0; }
        else if (warp == 2) { nq = 1; qi[0] = 2; qj[0] = 1; qi[1] = 2; qj[1] = 1; }
        else { nq = 1; qi[0] = 2; qj[0] = 2; qi[1] = 2; qj[1] = 2; }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (q < nq) {
                const int i0 = 8 * qi[q], j0 = 8 * qj[q];
                const int i = i0 + g, j = j0 + 2 * t;
                const double2 y2 = *reinterpret_cast<const double2*>(v.Y + i * 24 + j);
                cq[q][0] = lxx_entry(i, j, cm, dt, dt, false) + y2.x + ((i == j) ? reg : 0.0);
                cq[q][1] = lxx_entry(i, j + 1, cm, dt, dt, false) + y2.y + ((i == j + 1) ? reg : 0.0);
#pragma unroll
                for (int kk = 0; kk < 12; kk += 4) dmma884(cq[q], v.At12[(kk + t) * 24 + i0 + g], v.Y[(kk + t) * 24 + j0 + g]);
            }
        }
        // Qux_r tiles (Ic, J): Ic in {0,1}, J in {0,1,2}; Quu_r tiles (Ic, Jc) in {0,1}^2
        {
            // job list per warp: warp0: Qux(0,0),Qux(1,0) ; warp1: Qux(0,1),Qux(1,1) ; warp2: Qux(0,2),Qux(1,2),Quu(0,0) ; warp3: Quu(0,1),Quu(1,0),Quu(1,1)
            const int njobs = (warp < 2) ? 2 : 3;
#pragma unroll
            for (int job = 0; job < 3; ++job) {
                if (job < njobs) {
                    bool is_quu;
                    int Ic, J;
                    if (warp < 2) { is_quu = false; Ic = job; J = warp; }
                    else if (warp == 2) { is_quu = (job == 2); Ic = is_quu ? 0 : job; J = is_quu ? 0 : 2; }
                    else { is_quu = true; Ic = (job == 0) ? 0 : 1; J = (job == 1) ? 0 : 1; }
                    const double* Bop = is_quu ? v.Zr : v.Y;
                    double cc[2] = {0.0, 0.0};
#pragma unroll
                    for (int kk = 4; kk < 12; kk += 4) dmma884(cc, v.Bq[(kk - 4 + t) * 24 + 8 * Ic + g], Bop[(kk + t) * 24 + 8 * J + g]);
                    // swing rows: (B_r^T M)[c][:] = dt * M[12+c][:]
                    const int c = 8 * Ic + g;
                    if (c < 12 && !((cm >> (c / 3)) & 1u)) {
                        const double2 m2 = *reinterpret_cast<const double2*>(Bop + (12 + c) * 24 + 8 * J + 2 * t);
                        cc[0] += pc.swdt[c / 3] * m2.x;
                        cc[1] += pc.swdt[c / 3] * m2.y;
                    }
                    if (is_quu) {  // + luu_r: dt R + reg on the diagonal, ReB blocks for stance legs
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int c2 = 8 * J + 2 * t + q;
                            if (c < 12 && c2 < 12) {
                                if (c == c2) cc[q] += dt * weight_R(act_index(c, cm)) + reg;
                                if (c / 3 == c2 / 3 && ((cm >> (c / 3)) & 1u)) cc[q] += sm.lq[LQ_LUU + 9 * (c / 3) + 3 * (c % 3) + (c2 % 3)];
                            }
                        }
                        *reinterpret_cast<double2*>(v.QuuR + c * 24 + 8 * J + 2 * t) = make_double2(cc[0], cc[1]);
                    } else {
                        *reinterpret_cast<double2*>(v.QuxR + c * 24 + 8 * J + 2 * t) = make_double2(cc[0], cc[1]);
                    }
                }
            }
        }
        if (warp == 3 && lane < 24) {  // Qx = lx + A^T Gn
            double acc = sm.Gn[lane];
#pragma unroll
            for (int r = 0; r < 9; ++r) acc = fma(v.At12[r * 24 + lane], sm.Gn[r], acc);
            sm.Qx[lane] = sm.lq[LQ_LX + lane] + acc;
        }
        if (warp == 2 && lane < 12) {  // Qu_r = lu_r + B_r^T Gn
            const int c = lane;
            double acc;
            if ((cm >> (c / 3)) & 1u) {
                acc = 0.0;
#pragma unroll
                for (int r = 4; r < 12; ++r) acc = fma(v.Bq[(r - 4) * 24 + c], sm.Gn[r], acc);
            } else {
                acc = pc.swdt[c / 3] * sm.Gn[12 + c];
            }
            sm.Qu[c] = sm.lq[LQ_LU + act_index(c, cm)] + acc;
        }
        __syncthreads();
        // ---- P3: Gauss-Jordan tableau (warps 0,1), shifted PD test (warp 2), inactive controls (warp 3) ----
        if (warp < 3) {
            double col[12];
            // lanes 0..11: columns of Quu_r ; warp0 lanes 12..31: Qux_r columns 0..19 ; warp1 lanes 12..15: Qux_r 20..23, lane 16: Qu_r
            const double* src = nullptr;
            int stride = 24;
            if (lane < 12) src = v.QuuR + lane;
            else if (warp == 0) src = v.QuxR + (lane - 12);
            else if (warp == 1 && lane < 16) src = v.QuxR + (lane + 8);
            else if (warp == 1 && lane == 16) { src = sm.Qu; stride = 1; }
#pragma unroll
            for (int r = 0; r < 12; ++r) col[r] = src ? src[r * stride] : ((r == (lane % 12)) ? 1.0 : 0.0);
            if (warp == 2 && lane < 12) col[lane] -= 1e-9;  // Quu - 1e-9 I (Q7)
            const bool ok = gauss_jordan12(col, sm.red + 16 * warp);
            if (warp == 2) {
                if (lane == 0) sm.ibuf[0] = ok ? 1 : 0;
            } else if (lane >= 12) {
                if (warp == 0) {
                    const int j = lane - 12;
                    double* Kg = sm.K + (size_t)s * 288;
#pragma unroll
                    for (int r = 0; r < 12; ++r) { v.KrS[r * 24 + j] = -col[r]; Kg[r * 24 + j] = -col[r]; }
                } else if (lane < 16) {
                    const int j = lane + 8;
                    double* Kg = sm.K + (size_t)s * 288;
#pragma unroll
                    for (int r = 0; r < 12; ++r) { v.KrS[r * 24 + j] = -col[r]; Kg[r * 24 + j] = -col[r]; }
                } else if (lane == 16) {
                    double dvk = 0.0;
#pragma unroll
                    for (int r = 0; r < 12; ++r) {
                        sm.wu[r] = -col[r];                                   // dU_r
                        sm.dU[24 * s + act_index(r, cm)] = -col[r];
                        dvk = fma(sm.Qu[r], col[r], dvk);                     // -Qu^T dU
                    }
                    sm.dbuf[0] = dvk;
                }
            }
        } else if (lane < 16) {
            // decoupled controls: Quu_ii = dt R_i + reg, Qu_i = lu_i, K row = 0
            double dv = 0.0;
            if (lane < 12) {
                const int i = inact_index(lane, cm);
                const double qu = sm.lq[LQ_LU + i];
                const double du = -qu / (dt * weight_R(i) + reg);
                sm.dU[24 * s + i] = du;
                dv = -qu * du;
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) dv += __shfl_xor_sync(0x0000ffffu, dv, o, 16);
            if (lane == 0) sm.dbuf[1] = dv;
        }
        __syncthreads();
        if (!sm.ibuf[0]) return false;
        // ---- P4: H' = sym(Qxx) + Qux_r^T K_r ; G' = Qx + Qux_r^T dU_r ----
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (q < nq) {
                const int i0 = 8 * qi[q], j0 = 8 * qj[q];
#pragma unroll
                for (int kk = 0; kk < 12; kk += 4) dmma884(cq[q], v.QuxR[(kk + t) * 24 + i0 + g], v.KrS[(kk + t) * 24 + j0 + g]);
                if (i0 == j0) {
                    // symmetrise the diagonal tile: partner of (g, 2t+q') is (2t+q', g), held by lane 4*(2t+q') + g/2, slot g&1
                    const double p00 = __shfl_sync(0xffffffffu, cq[q][0], 4 * (2 * t) + (g >> 1));
                    const double p01 = __shfl_sync(0xffffffffu, cq[q][1], 4 * (2 * t) + (g >> 1));
                    const double p10 = __shfl_sync(0xffffffffu, cq[q][0], 4 * (2 * t + 1) + (g >> 1));
                    const double p11 = __shfl_sync(0xffffffffu, cq[q][1], 4 * (2 * t + 1) + (g >> 1));
                    cq[q][0] = 0.5 * (cq[q][0] + ((g & 1) ? p01 : p00));
                    cq[q][1] = 0.5 * (cq[q][1] + ((g & 1) ? p11 : p10));
                    *reinterpret_cast<double2*>(v.H + (i0 + g) * 24 + j0 + 2 * t) = make_double2(cq[q][0], cq[q][1]);
                } else {
                    *reinterpret_cast<double2*>(v.H + (i0 + g) * 24 + j0 + 2 * t) = make_double2(cq[q][0], cq[q][1]);
                    v.H[(j0 + 2 * t) * 24 + i0 + g] = cq[q][0];
                    v.H[(j0 + 2 * t + 1) * 24 + i0 + g] = cq[q][1];
                }
            }
        }
        if (warp == 3 && lane < 24) {
            double acc = sm.Qx[lane];
#pragma unroll
            for (int r = 0; r < 12; ++r) acc = fma(v.QuxR[r * 24 + lane], sm.wu[r], acc);
            sm.G[lane] = acc;
        }
        const double dvk = sm.dbuf[0] + sm.dbuf[1];
        dV1 -= dvk;
        dV2 += dvk;
        __syncthreads();
    }
    // G[0] += H[0] * Defect[0]
    {
        const int n0 = sc.node_off[ph];
        if (tid < 24) sm.dfc[tid] = sm.Defect[24 * n0 + tid];
        __syncthreads();
        double acc = 0.0;
        if (tid < 24) {
#pragma unroll
            for (int j = 0; j < 24; ++j) acc = fma(v.H[tid * 24 + j], sm.dfc[j], acc);
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
            const double* P = sm.Y;  // column-major P[r + 24 c]
            for (int e = tid; e < 576; e += kThreads) {  // Z = P^T H  (row-major Z[i][j])
                const int i = e / 24, j = e % 24;
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < 24; ++m) acc = fma(P[m + 24 * i], sm.H[m * 24 + j], acc);
                sm.Z[e] = acc;
            }
            if (tid < 24) {
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < 24; ++m) acc = fma(P[m + 24 * tid], sm.G[m], acc);
                sm.vtmp[tid] = acc;
            }
            __syncthreads();
            for (int e = tid; e < 576; e += kThreads) {  // H = Z P
                const int i = e / 24, j = e % 24;
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < 24; ++m) acc = fma(sm.Z[i * 24 + m], P[m + 24 * j], acc);
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
        for (int e = tid; e < 576; e += kThreads) sm.g0h0[24 + e] = sm.H[e];  // symmetric: row-major == column-major
        if (tid < 24) sm.g0h0[tid] = sm.G[tid];
    }
    __syncthreads();
    if (tid == 0) { sm.st.dV_1 = dV1; sm.st.dV_2 = dV2; sm.st.sweep_ok = success ? 1 : 0; }
    __syncthreads();
    return success;
}

// MultiPhaseDDP::backward_sweep_regularized (Q8).  Returns success; counts sweeps.
__device__ inline bool backward_sweep_regularized_block(Smem& sm, int& n_sweeps) {
    bool success = false;
    double reg = sm.st.reg;
    n_sweeps = 0;
    __syncthreads();
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
    double* sdu = sm.vtmp2;  // current du (full 24)
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
                for (int m = 0; m < 24; ++m) acc = fma(sm.Y[tid + 24 * m], sdx[m], acc);
                sm.Gn[tid] = acc;
            }
            __syncthreads();
        } else {
            if (tid < 24) sm.Gn[tid] = 0.0;
            __syncthreads();
        }
        if (tid < 32) {
            // lane -> reduced control row it owns in K_r (lanes 0..11), and its full control index
            const bool stance_of_lane = (lane < 24) && ((cm >> ((lane % 12) / 3)) & 1u);
            const bool lane_active = (lane < 24) && ((lane < 12) == stance_of_lane);  // control `lane` is coupled
            const int cred = lane % 12;
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
                    double acc = 0.0;
                    if (lane_active) {
                        const double* Kr = sm.K + (size_t)s * 288 + cred * 24;
#pragma unroll 8
                        for (int j = 0; j < 24; ++j) acc = fma(Kr[j], sdx[j], acc);
                    }
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
                        for (int j = 0; j < 24; ++j) acc = fma(At[r * 24 + j], sdx[j], acc);
                        adx += acc;
                    } else if (lane >= 3 && lane < 6) {
                        adx += dt * sdx[lane + 6];
                    }
                    // B du
                    double bdu;
                    if (lane >= 6 && lane < 9) {
                        double acc = 0.0;
#pragma unroll
                        for (int j = 0; j < 12; ++j) acc = fma(Bt[(lane - 6) * 12 + j], sdu[j], acc);
                        bdu = acc;
                    } else if (lane >= 9 && lane < 12) {
                        double acc = 0.0;
#pragma unroll
                        for (int l = 0; l < 4; ++l) acc = fma(pc.cm[l], sdu[3 * l + lane - 9], acc);
                        bdu = acc;
                    } else if (lane >= 12) {
                        bdu = pc.swdt[(lane - 12) / 3] * du;
                    } else {
                        bdu = 0.0;
                    }
                    dxn = (adx + bdu) + eps * sm.Defect[24 * (n0 + k + 1) + lane];
                    // expected cost change, lane-partial
                    dV1 += rec[LQ_LX + lane] * dx + rec[LQ_LU + lane] * du;
                    double qdx = lxx_entry(lane, lane, cm, dt, dt, false) * dx;
                    if (lane >= 3 && lane < 6) {
                        for (int l = 0; l < 4; ++l) qdx += lxx_entry(lane, 12 + 3 * l + lane - 3, cm, dt, dt, false) * sdx[12 + 3 * l + lane - 3];
                    } else if (lane >= 12) {
                        const int jj = (lane - 12) % 3;
                        qdx += lxx_entry(lane, 3 + jj, cm, dt, dt, false) * sdx[3 + jj];
                    }
                    dV2 += dx * qdx;
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
