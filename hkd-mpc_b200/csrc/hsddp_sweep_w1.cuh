// Backward sweep, one WARP per problem (the phased driver's sweep kernel for large batches).
//
//   MultiPhaseDDP::backward_sweep_regularized   HSDDPSolver/source/MultiPhaseDDP.cpp:141-181
//   MultiPhaseDDP::backward_sweep               :190-229  (+ impact_aware_step :480-484)
//   SinglePhase::backward_sweep                 HSDDPSolver/source/SinglePhase.cpp:299-367
//
// Same algebra, tile decomposition and accumulation order as phase_backward_sweep_block (hsddp_sweep.cuh: one 4-warp
// block per problem), but the whole stage runs in ONE warp:
//   * no block barriers: the four barriers per stage of the block version cost 32 % of its issue slots
//     (profiles/r02a: barrier 3.9 of 12 stalled warps per issue) and half of its warps idle during the elimination;
//   * the tensor-core operands of a phase are loaded ONCE into registers and shared by all its output tiles
//     (P1: 9 + 15 operand loads for 15 tiles instead of 60; P2: 30 for 16 tiles; P4: 18 for 6), and the tiles of a
//     group are issued interleaved (independent accumulators), so one warp keeps its sub-partition's FP64 tensor pipe
//     busy: shared-memory wavefronts per stage drop from ~1,460 to ~800 (the block version ran at 64 % of the
//     shared-memory pipe, its hard floor);
//   * the 12x12 block Gauss-Jordan carries TWO tableau columns per lane (49 columns: 12 Quu_r, 24 Qux_r, Qu_r, 12
//     identity), so the pivot columns are published and read once instead of once per eliminating warp.
// Eight problems (warps) are resident per SM (27.7 KB of shared memory and 254 registers each): while one warp sits in the
// latency-bound elimination the other warp of its sub-partition owns the tensor pipe.
#pragma once
#include "hsddp_sweep.cuh"

namespace hsddp {

struct __align__(16) SweepW1 {
    double H[ro(24)], Y[ro(24)];
    double Z[zo(24)];               // H B_r [24][16]; after P2: K_r^T [24][12]
    double Qux[ro(12) + kQuuPad];   // Qux_r [12][24] (rows 12..15 of its second row block are never stored); with Quu it is also
    double Quu[ro(12)];             // the 24-row temporary of the impact-aware step
    double R[hkd::kRSize];          // dense [A - I | B_r] rows 0..11, [12][44]
    double cr[CR_STRIDE];           // the compact record of the stage (HBM layout): entries for R, then lx, lu, luu; refilled by
    double dfc[24];                 // cp.async for the NEXT stage once P2 has consumed lx, lu, luu (defect: once P1 has)
    double zvec[24];                // 0 x 11, 1, 0 x 12: the identity / zero tableau columns are slices of it
    double G[24], Gn[24], Qx[24], vtmp[24];
    double Qu[12], wu[12];
    double lxxd[24], lxxTd[24], lxxw[12], lxxTw[12];
    double swc[16], swdt[4];
    double sbuf[24];                // the two pivot columns of an elimination step
    double dbuf[4];
    int n_phases, n_stages, verdict, _pad;
    int horizon[MAXPH], node_off[MAXPH], stage_off[MAXPH];
    unsigned cmask[MAXPH], nmask[MAXPH];
};
static_assert(offsetof(SweepW1, Quu) - offsetof(SweepW1, Qux) == sizeof(double) * (ro(12) + kQuuPad), "Qux and Quu must be contiguous");
static_assert(ro(12) + kQuuPad + ro(12) >= ro(24), "Qux|Quu must hold a 24-row temporary");
static_assert(sizeof(SweepW1) + 1024 <= 233472 / 8, "eight problems per SM");

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double x, double y) { *reinterpret_cast<double2*>(p) = make_double2(x, y); }

// Block Gauss-Jordan with 2x2 pivots (gauss_jordan12 of hsddp_sweep.cuh) on two tableau columns per lane.  The pivot
// matrix lives in slot `va` of lanes 0..11.
__device__ __forceinline__ bool gauss_jordan12x2(double (&va)[12], double (&vb)[12], double* sbuf) {
    const int lane = threadIdx.x & 31;
    bool ok = true;
#pragma unroll 1
    for (int step = 0; step < 6; ++step) {
        if ((lane >> 1) == step) {
            double2* dst = reinterpret_cast<double2*>(sbuf + 12 * (lane & 1));
#pragma unroll
            for (int r = 0; r < 12; r += 2) dst[r >> 1] = make_double2(va[r], va[r + 1]);
        }
        __syncwarp();
        const double2 pk = ld2(sbuf), pk1 = ld2(sbuf + 12);
        const double det = pk.x * pk1.y - pk1.x * pk.y;
        if (pk.x < 0.0 || det < 0.0) ok = false;
        const double rdet = pivot_rcp(det);
        const double ta0 = (pk1.y * va[0] - pk1.x * va[1]) * rdet;
        const double ta1 = (pk.x * va[1] - pk.y * va[0]) * rdet;
        const double tb0 = (pk1.y * vb[0] - pk1.x * vb[1]) * rdet;
        const double tb1 = (pk.x * vb[1] - pk.y * vb[0]) * rdet;
#pragma unroll
        for (int r = 2; r < 12; r += 2) {
            const double2 a = ld2(sbuf + r), b = ld2(sbuf + 12 + r);
            va[r - 2] = fma(-b.x, ta1, fma(-a.x, ta0, va[r]));
            va[r - 1] = fma(-b.y, ta1, fma(-a.y, ta0, va[r + 1]));
            vb[r - 2] = fma(-b.x, tb1, fma(-a.x, tb0, vb[r]));
            vb[r - 1] = fma(-b.y, tb1, fma(-a.y, tb0, vb[r + 1]));
        }
        va[10] = ta0; va[11] = ta1;
        vb[10] = tb0; vb[11] = tb1;
        __syncwarp();
    }
    return ok;
}

struct SweepW1Ptrs {  // per-problem HBM pointers (registers, warp-uniform)
    const double *lqg, *tq, *Defect;
    double *K, *dU, *g0h0;
};

// fetch the compact record of stage s and the defect of node n1
__device__ __forceinline__ void w1_prefetch(SweepW1& sm, const SweepW1Ptrs& p, int s, int n1) {
    const int lane = threadIdx.x & 31;
    const char* src = reinterpret_cast<const char*>(p.lqg + (size_t)s * CR_STRIDE);
    char* dst = reinterpret_cast<char*>(sm.cr);
#pragma unroll
    for (int u = lane; u < 98; u += 32) cp_async16(dst + 16 * u, src + 16 * u);  // 196 doubles: entries, lx, lu, luu
    if (lane < 12) cp_async16(reinterpret_cast<char*>(sm.dfc) + 16 * lane, reinterpret_cast<const char*>(p.Defect + 24 * n1) + 16 * lane);
}

__device__ inline void w1_phase_tables(SweepW1& sm, unsigned cm, double dt) {
    const int lane = threadIdx.x & 31;
    if (lane < 24) {
        const int q = lane % 12, l = q / 3, jj = q % 3;
        const double c = (double)((cm >> l) & 1u);
        const double scale = (lane < 12) ? dt : 20.0;
        (lane < 12 ? sm.lxxw : sm.lxxTw)[q] = (scale * c * weight_foot(l, jj, cm)) * c;
    } else if (lane < 28) {
        const double c = (double)((cm >> (lane - 24)) & 1u);
        sm.swdt[lane - 24] = (1.0 - c) * dt;
    }
    if (lane < 16) sm.swc[lane] = (lane < 12) ? (1.0 - (double)((cm >> (lane / 3)) & 1u)) * dt : 0.0;
    __syncwarp();
#pragma unroll
    for (int e = lane; e < 48; e += 32) {
        const bool term = e >= 24;
        const int i = e % 24;
        const double* w = term ? sm.lxxTw : sm.lxxw;
        double val = term ? weight_Qf(i, cm) : dt * weight_Q(i, cm);
        if (i >= 3 && i < 6) { for (int l = 0; l < 4; ++l) val += w[3 * l + i - 3]; }
        else if (i >= 12) val += w[i - 12];
        (term ? sm.lxxTd : sm.lxxd)[i] = val;
    }
    __syncwarp();
}

// One phase of the backward sweep (SinglePhase::backward_sweep); sm.G / sm.H hold Gprime / Hprime on entry.
__device__ inline bool w1_phase_sweep(SweepW1& sm, const SweepW1Ptrs& p, int ph, double reg, double dt, const int (&rpos)[4], double& dV1, double& dV2) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const unsigned cm = sm.cmask[ph];
    const int Nph = sm.horizon[ph];
    const double* trec = p.tq + ph * TQ_STRIDE;
    w1_phase_tables(sm, cm, dt);
    w1_prefetch(sm, p, sm.stage_off[ph] + Nph - 1, sm.node_off[ph] + Nph);
    // fragment base pointers: accumulator element (g, 2t..2t+1), operand element (t, g)
    const double* hA = sm.H + ro(t) + g;
    double* hC = sm.H + ro(g) + 2 * t;
    double* hT = sm.H + ro(2 * t) + g;
    double* yC = sm.Y + ro(g) + 2 * t;
    const double* yB = sm.Y + ro(t) + g;
    double* zC = sm.Z + zo(g) + 2 * t;
    const double* zB = sm.Z + zo(t) + g;
    const double* qA = sm.Qux + ro(t) + g;
    const double* kB = sm.Z + g * 12 + t;
    double* quxC = sm.Qux + ro(g) + 2 * t;
    double* quuC = sm.Quu + ro(g) + 2 * t;
    const double* rB = sm.R + t * hkd::kRld + g;
    // G[N] = Phix + Gprime ; H[N] = Phixx + Hprime (sparse Phixx: diagonal, foot-regulariser coupling, AL outer products)
    if (lane < 24) {
        sm.G[lane] += trec[TQ_PHIX + lane];
        sm.H[ro(lane) + lane] += sm.lxxTd[lane];
    }
    if (lane < 24) {
        const int q = lane % 12, j3 = 3 + q % 3;
        sm.H[(lane < 12) ? ro(12 + q) + j3 : ro(j3) + 12 + q] -= sm.lxxTw[q];
    }
    __syncwarp();
#pragma unroll 1
    for (int e = lane; e < 49; e += 32) {
        const int a = e / 7, b = e % 7;
#pragma unroll 1
        for (int l = 0; l < 4; ++l) {
            const double wh = trec[TQ_WH + l];
            if (wh != 0.0) {
                const int i = (a < 3) ? a : (a == 3) ? 5 : 12 + 3 * l + a - 4;
                const int j = (b < 3) ? b : (b == 3) ? 5 : 12 + 3 * l + b - 4;
                sm.H[ro(i) + j] += wh * (trec[TQ_HX + 24 * l + i] * trec[TQ_HX + 24 * l + j]);
            }
        }
    }
    dV1 = 0.0; dV2 = 0.0;
#pragma unroll 1
    for (int k = Nph - 1; k >= 0; --k) {
        const int s = sm.stage_off[ph] + k;
        cp_async_wait_all();
        __syncwarp();
        {   // compact record -> dense tile (the other entries of the tile are constant)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = lane + 32 * q;
                if (i < hkd::kCrNnz) sm.R[rpos[q]] = sm.cr[CR_R + i];
            }
        }
        const double* lxv = sm.cr + CR_LX;
        const double* luv = sm.cr + CR_LU;
        const double* luu = sm.cr + CR_LUU;
        const double* dfc = sm.dfc;
        __syncwarp();
        // ---- P1: [Y | Z] = H [A | B_r], Gn = G + H d ----
        {
            double a[3][3];  // H[:, 0..11] operand fragments: row block I, k block
#pragma unroll
            for (int I = 0; I < 3; ++I) { a[I][0] = hA[8 * I]; a[I][1] = hA[8 * I + RO4]; a[I][2] = hA[8 * I + RO8]; }
#pragma unroll
            for (int J = 0; J < 3; ++J) {  // Y = H + H At : column block J, three row blocks interleaved
                const double b0 = rB[8 * J], b1 = rB[8 * J + 4 * hkd::kRld], b2 = rB[8 * J + 8 * hkd::kRld];
                double c[3][2];
#pragma unroll
                for (int I = 0; I < 3; ++I) { const double2 h2 = ld2(hC + ro(8 * I) + 8 * J); c[I][0] = h2.x; c[I][1] = h2.y; }
#pragma unroll
                for (int I = 0; I < 3; ++I) dmma(c[I], a[I][0], b0);
#pragma unroll
                for (int I = 0; I < 3; ++I) dmma(c[I], a[I][1], b1);
#pragma unroll
                for (int I = 0; I < 3; ++I) dmma(c[I], a[I][2], b2);
#pragma unroll
                for (int I = 0; I < 3; ++I) st2(yC + ro(8 * I) + 8 * J, c[I][0], c[I][1]);
            }
#pragma unroll
            for (int Jz = 0; Jz < 2; ++Jz) {  // Z = H B_r (+ swing columns: H[:, 12+c] (1-c_l) dt)
                const double b0 = rB[24 + 8 * Jz], b1 = rB[24 + 8 * Jz + 4 * hkd::kRld], b2 = rB[24 + 8 * Jz + 8 * hkd::kRld];
                const double2 sw = ld2(sm.swc + 2 * t + 8 * Jz);
                double c[3][2];
#pragma unroll
                for (int I = 0; I < 3; ++I) { c[I][0] = 0.0; c[I][1] = 0.0; }
#pragma unroll
                for (int I = 0; I < 3; ++I) dmma(c[I], a[I][0], b0);
#pragma unroll
                for (int I = 0; I < 3; ++I) dmma(c[I], a[I][1], b1);
#pragma unroll
                for (int I = 0; I < 3; ++I) dmma(c[I], a[I][2], b2);
#pragma unroll
                for (int I = 0; I < 3; ++I) {
                    const double2 h2 = ld2(hC + ro(8 * I) + 12 + 8 * Jz);
                    st2(zC + zo(8 * I) + 8 * Jz, fma(h2.x, sw.x, c[I][0]), fma(h2.y, sw.y, c[I][1]));
                }
            }
            if (lane < 24) {  // Gn = G + H d   (Q10)
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int j = 0; j < 24; j += 4) {
                    a0 = fma(sm.H[ro(j) + lane], dfc[j], a0);
                    a1 = fma(sm.H[ro(j + 1) + lane], dfc[j + 1], a1);
                    a2 = fma(sm.H[ro(j + 2) + lane], dfc[j + 2], a2);
                    a3 = fma(sm.H[ro(j + 3) + lane], dfc[j + 3], a3);
                }
                sm.Gn[lane] = sm.G[lane] + ((a0 + a1) + (a2 + a3));
            }
        }
        __syncwarp();
        // ---- P2: Qux_r = B_r^T Y, Quu_r = luu_r + B_r^T Z, Qxx = Y + At^T Y (lower tiles, parked in H), Qx, Qu_r ----
        {
            double bY[3][3];  // Y[0..11][:] operand fragments: column block J, k block
#pragma unroll
            for (int J = 0; J < 3; ++J) { bY[J][0] = yB[8 * J]; bY[J][1] = yB[8 * J + RO4]; bY[J][2] = yB[8 * J + RO8]; }
            double aB[2][3];  // B_r^T operand fragments: row block Ci, k block
#pragma unroll
            for (int Ci = 0; Ci < 2; ++Ci) { aB[Ci][0] = rB[24 + 8 * Ci]; aB[Ci][1] = rB[24 + 8 * Ci + 4 * hkd::kRld]; aB[Ci][2] = rB[24 + 8 * Ci + 8 * hkd::kRld]; }
#pragma unroll
            for (int Ci = 0; Ci < 2; ++Ci) {  // Qux_r rows 8 Ci + g
                const int c = 8 * Ci + g;
                const double sw = (c < 12) ? sm.swc[c] : 0.0;
                double acc[3][2];
#pragma unroll
                for (int J = 0; J < 3; ++J) { acc[J][0] = 0.0; acc[J][1] = 0.0; }
#pragma unroll
                for (int J = 0; J < 3; ++J) dmma(acc[J], aB[Ci][0], bY[J][0]);
#pragma unroll
                for (int J = 0; J < 3; ++J) dmma(acc[J], aB[Ci][1], bY[J][1]);
#pragma unroll
                for (int J = 0; J < 3; ++J) dmma(acc[J], aB[Ci][2], bY[J][2]);
                if (c < 12) {
#pragma unroll
                    for (int J = 0; J < 3; ++J) {
                        if (sw != 0.0) {  // swing row: (B_r^T Y)[c][:] = (1-c_l) dt * Y[12+c][:]
                            const double2 m2 = ld2(sm.Y + ro(12 + c) + 8 * J + 2 * t);
                            acc[J][0] = fma(sw, m2.x, acc[J][0]);
                            acc[J][1] = fma(sw, m2.y, acc[J][1]);
                        }
                        st2(quxC + ro(8 * Ci) + 8 * J, acc[J][0], acc[J][1]);
                    }
                }
            }
            {   // Quu_r: four tiles (Ci, J)
                double bZ[2][3];
#pragma unroll
                for (int J = 0; J < 2; ++J) { bZ[J][0] = zB[8 * J]; bZ[J][1] = zB[8 * J + ZO4]; bZ[J][2] = zB[8 * J + ZO8]; }
                double acc[2][2][2];
#pragma unroll
                for (int Ci = 0; Ci < 2; ++Ci)
#pragma unroll
                    for (int J = 0; J < 2; ++J) { acc[Ci][J][0] = 0.0; acc[Ci][J][1] = 0.0; }
#pragma unroll
                for (int kb = 0; kb < 3; ++kb)
#pragma unroll
                    for (int Ci = 0; Ci < 2; ++Ci)
#pragma unroll
                        for (int J = 0; J < 2; ++J) dmma(acc[Ci][J], aB[Ci][kb], bZ[J][kb]);
#pragma unroll
                for (int Ci = 0; Ci < 2; ++Ci) {
                    const int c = 8 * Ci + g;
                    if (c < 12) {
                        const double sw = sm.swc[c];
                        const bool stance = sw == 0.0;
                        const int l3 = 3 * (c / 3);
                        const double diag = dt * (stance ? .2 : .1) + reg;  // weight_R(act_index(c, cm))
                        const double* lb = luu + 3 * c;                      // luu[9 (c/3) + 3 (c%3) + k]
#pragma unroll
                        for (int J = 0; J < 2; ++J) {
                            double c0 = acc[Ci][J][0], c1 = acc[Ci][J][1];
                            if (sw != 0.0) {  // swing row: (B_r^T Z)[c][:] = (1-c_l) dt * Z[12+c][:]
                                const double2 m2 = ld2(sm.Z + zo(12 + c) + 8 * J + 2 * t);
                                c0 = fma(sw, m2.x, c0);
                                c1 = fma(sw, m2.y, c1);
                            }
                            const int cc = 8 * J + 2 * t;
                            if (c == cc) c0 += diag;
                            if (c == cc + 1) c1 += diag;
                            if (stance) {
                                if (cc >= l3 && cc < l3 + 3) c0 += lb[cc - l3];
                                if (cc + 1 >= l3 && cc + 1 < l3 + 3) c1 += lb[cc + 1 - l3];
                            }
                            st2(quuC + ro(8 * Ci) + 8 * J, c0, c1);
                        }
                    }
                }
            }
            {   // Qxx lower tiles (I, J): Y tile + At^T[:, I-block]^T Y[:, J-block]
                double aA[3][3];
#pragma unroll
                for (int I = 0; I < 3; ++I) { aA[I][0] = rB[8 * I]; aA[I][1] = rB[8 * I + 4 * hkd::kRld]; aA[I][2] = rB[8 * I + 8 * hkd::kRld]; }
                double c[6][2];
                constexpr int TI[6] = {0, 1, 2, 1, 2, 2}, TJ[6] = {0, 0, 0, 1, 1, 2};
#pragma unroll
                for (int q = 0; q < 6; ++q) { const double2 y2 = ld2(yC + ro(8 * TI[q]) + 8 * TJ[q]); c[q][0] = y2.x; c[q][1] = y2.y; }
#pragma unroll
                for (int kb = 0; kb < 3; ++kb)
#pragma unroll
                    for (int q = 0; q < 6; ++q) dmma(c[q], aA[TI[q]][kb], bY[TJ[q]][kb]);
                __syncwarp();  // (every lane has read its H operands of P1 / Gn before Qxx overwrites H)
#pragma unroll
                for (int q = 0; q < 6; ++q) st2(hC + ro(8 * TI[q]) + 8 * TJ[q], c[q][0], c[q][1]);
            }
            if (lane < 24) {  // Qx = lx + A^T Gn
                double acc = sm.Gn[lane];
#pragma unroll
                for (int r = 0; r < 9; ++r) acc = fma(sm.R[r * hkd::kRld + lane], sm.Gn[r], acc);
                sm.Qx[lane] = lxv[lane] + acc;
            }
            if (lane < 12) {  // Qu_r = lu_r + B_r^T Gn
                const int c = lane;
                double acc = 0.0;
                if ((cm >> (c / 3)) & 1u) {
#pragma unroll
                    for (int r = 6; r < 12; ++r) acc = fma(sm.R[r * hkd::kRld + 24 + c], sm.Gn[r], acc);
                } else {
                    acc = sm.swdt[c / 3] * sm.Gn[12 + c];
                }
                sm.Qu[c] = luv[act_index(c, cm)] + acc;
            }
            __syncwarp();
            // sparse additive part of Qxx: lxx + reg I on the diagonal, foot-regulariser coupling of the lower triangle
            if (lane < 24) sm.H[ro(lane) + lane] = (sm.H[ro(lane) + lane] + sm.lxxd[lane]) + reg;
            if (lane < 12) sm.H[ro(12 + lane) + 3 + lane % 3] -= sm.lxxw[lane];
            if (lane >= 16) {  // decoupled controls (lanes 16..27): Quu_ii = dt R_i + reg, Qu_i = lu_i, K row = 0
                const int q = lane - 16;
                double dv = 0.0;
                if (q < 12) {
                    const int i = inact_index(q, cm);
                    const double qu = luv[i];
                    const double du = -qu / (dt * weight_R(i) + reg);
                    p.dU[24 * s + i] = du;
                    dv = -qu * du;
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) dv += __shfl_xor_sync(0xffff0000u, dv, o, 16);
                if (q == 0) sm.dbuf[1] = dv;
            }
        }
        __syncwarp();
        if (k > 0) w1_prefetch(sm, p, s - 1, sm.node_off[ph] + k);  // (lx, lu, luu and the defect of this stage have been consumed)
        // ---- P3: block Gauss-Jordan on the tableau [Quu_r | Qux_r | Qu_r | I], two columns per lane ----
        // slot a: lanes 0..11 Quu_r columns, lanes 12..31 Qux_r columns 0..19
        // slot b: lanes 0..3 Qux_r columns 20..23, lane 4 Qu_r, lanes 5..16 identity columns, lanes 17..31 zero
        // PD verdict (LDLT(Quu - 1e-9 I).isPositive(), Q7) as in the block version: non-positive pivot => false; all pivots
        // positive and every column of Quu_r^-1 shorter than 5e8 / sqrt(12) => true; otherwise an exact second pass.
        int verdict;
#pragma unroll 1
        for (int pass = 0;; ++pass) {
            double va[12], vb[12];
            const bool a_piv = lane < 12;
            const int ja = lane - 12;  // Qux_r column of slot a (lanes >= 12)
            {
                const double* src = a_piv ? sm.Quu + lane : sm.Qux + ja;
#pragma unroll
                for (int r = 0; r < 12; ++r) va[r] = src[ro(r)];
                if (pass && a_piv) {
#pragma unroll
                    for (int r = 0; r < 12; ++r) if (r == lane) va[r] -= 1e-9;
                }
            }
            const bool b_gain = lane < 4, b_ff = lane == 4, b_inv = lane >= 5 && lane < 17;
            {   // slot b: a strided Qux_r column, or a unit-stride slice: Qu_r, column lane - 5 of the identity (zvec + 11 - (lane - 5)), zeros
                const double* src = b_gain ? sm.Qux + 20 + lane : b_ff ? sm.Qu : b_inv ? sm.zvec + 16 - lane : sm.zvec + 12;
#pragma unroll
                for (int r = 0; r < 12; ++r) vb[r] = src[b_gain ? ro(r) : r];
            }
            const bool ok = gauss_jordan12x2(va, vb, sm.sbuf);
            if (pass) { verdict = ok ? 1 : 0; break; }
            // gains: K_r[:, j] = -Quu_r^-1 Qux_r[:, j] -> KT[j][0..11] (smem + HBM)
            if (!a_piv) {
                double2* ks = reinterpret_cast<double2*>(sm.Z + 12 * ja);
                double2* kg = reinterpret_cast<double2*>(p.K + (size_t)s * 288 + 12 * ja);
#pragma unroll
                for (int r = 0; r < 12; r += 2) { const double2 v = make_double2(-va[r], -va[r + 1]); ks[r >> 1] = v; kg[r >> 1] = v; }
            }
            if (b_gain) {
                double2* ks = reinterpret_cast<double2*>(sm.Z + 12 * (20 + lane));
                double2* kg = reinterpret_cast<double2*>(p.K + (size_t)s * 288 + 12 * (20 + lane));
#pragma unroll
                for (int r = 0; r < 12; r += 2) { const double2 v = make_double2(-vb[r], -vb[r + 1]); ks[r >> 1] = v; kg[r >> 1] = v; }
            }
            // lane 4: Qu^T (Quu_r^-1 Qu_r) = -Qu^T dU ; lanes 5..16: squared norm of their column of Quu_r^-1
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int r = 0; r < 12; r += 4) {
                a0 = fma(b_ff ? sm.Qu[r] : vb[r], vb[r], a0);
                a1 = fma(b_ff ? sm.Qu[r + 1] : vb[r + 1], vb[r + 1], a1);
                a2 = fma(b_ff ? sm.Qu[r + 2] : vb[r + 2], vb[r + 2], a2);
                a3 = fma(b_ff ? sm.Qu[r + 3] : vb[r + 3], vb[r + 3], a3);
            }
            const double dot = (a0 + a1) + (a2 + a3);
            if (b_ff) {
#pragma unroll
                for (int r = 0; r < 12; ++r) {
                    sm.wu[r] = -vb[r];                                  // dU_r
                    p.dU[24 * s + act_index(r, cm)] = -vb[r];
                }
                sm.dbuf[0] = dot;
            }
            const bool big = __any_sync(0xffffffffu, b_inv && !(dot < 0.25e18 / 12));
            const bool all_ok = __all_sync(0xffffffffu, ok);
            verdict = !all_ok ? 0 : big ? 2 : 1;
            if (verdict != 2) break;
        }
        if (!verdict) { cp_async_wait_all(); __syncwarp(); return false; }
        __syncwarp();
        // ---- P4: H' = sym(Qxx) + Qux_r^T K_r (lower tiles, mirrored) ; G' = Qx + Qux_r^T dU_r ----
        {
            double a[3][3], b[3][3];  // Qux_r^T row block I / K_r^T column block J, k block
#pragma unroll
            for (int I = 0; I < 3; ++I) {
                a[I][0] = qA[8 * I]; a[I][1] = qA[8 * I + RO4]; a[I][2] = qA[8 * I + RO8];
                b[I][0] = kB[96 * I]; b[I][1] = kB[96 * I + 4]; b[I][2] = kB[96 * I + 8];
            }
            double c[6][2];
            constexpr int TI[6] = {0, 1, 2, 2, 1, 2}, TJ[6] = {0, 1, 2, 1, 0, 0};  // three diagonal tiles, then (2,1), (1,0), (2,0)
#pragma unroll
            for (int q = 0; q < 6; ++q) { const double2 q2 = ld2(hC + ro(8 * TI[q]) + 8 * TJ[q]); c[q][0] = q2.x; c[q][1] = q2.y; }
            double gp = 0.0;
            if (lane < 24) {  // G' = Qx + Qux_r^T dU_r
                double a0 = sm.Qx[lane], a1 = 0.0;
#pragma unroll
                for (int r = 0; r < 12; r += 2) {
                    a0 = fma(sm.Qux[ro(r) + lane], sm.wu[r], a0);
                    a1 = fma(sm.Qux[ro(r + 1) + lane], sm.wu[r + 1], a1);
                }
                gp = a0 + a1;
            }
#pragma unroll
            for (int kb = 0; kb < 3; ++kb)
#pragma unroll
                for (int q = 0; q < 6; ++q) dmma(c[q], a[TI[q]][kb], b[TJ[q]][kb]);
#pragma unroll
            for (int q = 0; q < 3; ++q) {  // symmetrise the diagonal tiles: partner of (g, 2t+q') is (2t+q', g), lane 4 (2t+q') + g/2, slot g & 1
                const double p00 = __shfl_sync(0xffffffffu, c[q][0], 4 * (2 * t) + (g >> 1));
                const double p01 = __shfl_sync(0xffffffffu, c[q][1], 4 * (2 * t) + (g >> 1));
                const double p10 = __shfl_sync(0xffffffffu, c[q][0], 4 * (2 * t + 1) + (g >> 1));
                const double p11 = __shfl_sync(0xffffffffu, c[q][1], 4 * (2 * t + 1) + (g >> 1));
                c[q][0] = 0.5 * (c[q][0] + ((g & 1) ? p01 : p00));
                c[q][1] = 0.5 * (c[q][1] + ((g & 1) ? p11 : p10));
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                st2(hC + ro(8 * TI[q]) + 8 * TJ[q], c[q][0], c[q][1]);
                if (q >= 3) {  // mirror (I, J) -> (J, I): rows 8 J + 2t, 2t + 1 ; column 8 I + g
                    hT[ro(8 * TJ[q]) + 8 * TI[q]] = c[q][0];
                    hT[ro(8 * TJ[q]) + 8 * TI[q] + 24] = c[q][1];  // row 2t+1 starts 24 after the even row 2t
                }
            }
            if (lane < 24) sm.G[lane] = gp;
        }
        const double dvk = sm.dbuf[0] + sm.dbuf[1];
        dV1 -= dvk;
        dV2 += dvk;
        __syncwarp();
    }
    // G[0] += H[0] * Defect[0]
    {
        const int n0 = sm.node_off[ph];
        if (lane < 24) sm.vtmp[lane] = p.Defect[24 * n0 + lane];
        __syncwarp();
        if (lane < 24) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 24; ++j) acc = fma(sm.H[ro(j) + lane], sm.vtmp[j], acc);  // H symmetric: conflict-free column read
            sm.G[lane] += acc;
        }
        __syncwarp();
    }
    return true;
}

// MultiPhaseDDP::backward_sweep(regularization): phases last -> first with the impact-aware step in between
__device__ inline bool w1_backward_sweep(SweepW1& sm, const SweepW1Ptrs& p, double reg, double dt, const int (&rpos)[4], double& dV1o, double& dV2o) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double dV1 = 0.0, dV2 = 0.0;
    bool success = true;
    for (int e = lane; e < hkd::kRSize; e += 32) {  // dense tile: zero + the constant dt of rows 3..5 of A - I
        const int r = e / hkd::kRld, c = e % hkd::kRld;
        sm.R[e] = (r >= 3 && r < 6 && c == 6 + r) ? dt : 0.0;
    }
    __syncwarp();
    for (int ph = sm.n_phases - 1; ph >= 0; --ph) {
        if (ph == sm.n_phases - 1) {
            for (int e = lane; e < ro(24); e += 32) sm.H[e] = 0.0;
            if (lane < 24) sm.G[lane] = 0.0;
            __syncwarp();
        } else {
            // impact-aware step: G' = Px^T G0, H' = Px^T H0 Px at the phase's terminal state (W = H P, then H' = P^T W)
            double* P = sm.Y;
            const double* Jc_all = p.tq + ph * TQ_STRIDE + TQ_JC;
            const unsigned c = sm.cmask[ph], cn = sm.nmask[ph];
            for (int e = lane; e < 576; e += 32) P[ro(e / 24) + e % 24] = ((e % 24) == (e / 24)) ? 1.0 : 0.0;
            __syncwarp();
            if (lane < 12) {  // HKDReset::resetmap_partial (HKDReset.h:78-136)
                const int l = lane / 3, r = lane % 3;
                const bool cl = (c >> l) & 1u, nl = (cn >> l) & 1u;
                const int row = 12 + 3 * l + r;
                double* Prow = P + ro(row);
                if (cl && !nl) Prow[row] = 0.0;
                if (!cl && nl) {
                    const double* Jc = Jc_all + 18 * l + 6 * r;
                    const double cmap = (r == 2) ? 0.0 : 1.0;
                    Prow[row] = 0.0;
                    for (int cc = 0; cc < 3; ++cc) {
                        Prow[cc] = cmap * Jc[cc];
                        Prow[3 + cc] = cmap * ((r == cc) ? 1.0 : 0.0);
                        Prow[12 + 3 * l + cc] = cmap * Jc[3 + cc];
                    }
                }
            }
            __syncwarp();
            double* W = sm.Qux;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const double* A = (pass ? P : sm.H) + ro(t) + g;
                const double* B = (pass ? W : P) + ro(t) + g;
                double* Cm = (pass ? sm.H : W) + ro(g) + 2 * t;
                double gacc = 0.0;
                if (!pass && lane < 24) {
#pragma unroll
                    for (int m = 0; m < 24; ++m) gacc = fma(P[ro(m) + lane], sm.G[m], gacc);
                }
#pragma unroll 1
                for (int J = 0; J < 3; ++J) {
                    double b[6];
#pragma unroll
                    for (int kk = 0; kk < 6; ++kk) b[kk] = B[8 * J + 104 * kk];  // ro(t + 4 kk) = ro(t) + 104 kk
                    double cc[3][2];
#pragma unroll
                    for (int I = 0; I < 3; ++I) { cc[I][0] = 0.0; cc[I][1] = 0.0; }
#pragma unroll
                    for (int kk = 0; kk < 6; ++kk)
#pragma unroll
                        for (int I = 0; I < 3; ++I) dmma(cc[I], A[8 * I + 104 * kk], b[kk]);
                    if (pass) __syncwarp();
#pragma unroll
                    for (int I = 0; I < 3; ++I) st2(Cm + ro(8 * I) + 8 * J, cc[I][0], cc[I][1]);
                }
                __syncwarp();
                if (!pass && lane < 24) sm.G[lane] = gacc;
            }
            __syncwarp();
        }
        double d1, d2;
        if (!w1_phase_sweep(sm, p, ph, reg, dt, rpos, d1, d2)) { success = false; break; }
        dV1 += d1;
        dV2 += d2;
    }
    if (success) {
        for (int e = lane; e < 576; e += 32) p.g0h0[24 + e] = sm.H[ro(e / 24) + e % 24];  // symmetric: row-major == column-major
        if (lane < 24) p.g0h0[lane] = sm.G[lane];
    }
    __syncwarp();
    dV1o = dV1; dV2o = dV2;
    return success;
}

// The sweep phase of one DDP iteration for every running problem (phased driver): backward_sweep_regularized with
// the bookkeeping of iter_sweep_block (hsddp_kernels.cu).  One warp = one block = one problem.
#ifndef HSDDP_W1_MINB
#define HSDDP_W1_MINB 8
#endif
__global__ void __launch_bounds__(32, HSDDP_W1_MINB) k_sweep_w1(BatchPtrs bp, hsddp_options opt) {
    __shared__ SweepW1 sm;
    const int lane = threadIdx.x;
    if (bp.n_active && ((int)blockIdx.x >= *bp.n_active || *bp.n_active < bp.sweep_w1_min)) return;  // (see BatchPtrs::active, sweep_w1_min)
    const int pid = bp.active ? bp.active[blockIdx.x] : (int)blockIdx.x;
    const DevSchedule* sc = bp.sched + bp.sched_id[pid];
    if (lane == 0) { sm.n_phases = sc->n_phases; sm.n_stages = sc->n_stages; }
    if (lane < 24) sm.zvec[lane] = (lane == 11) ? 1.0 : 0.0;
    if (lane < MAXPH) {
        sm.horizon[lane] = sc->horizon[lane]; sm.node_off[lane] = sc->node_off[lane]; sm.stage_off[lane] = sc->stage_off[lane];
        sm.cmask[lane] = sc->cmask[lane]; sm.nmask[lane] = sc->nmask[lane];
    }
    const double dt = sc->dt;
    SweepW1Ptrs p;
    p.lqg = bp.lq + (size_t)pid * bp.max_stages * CR_STRIDE;
    p.tq = bp.tq + (size_t)pid * MAXPH * TQ_STRIDE;
    p.Defect = bp.Defect + (size_t)pid * bp.max_nodes * 24;
    p.K = bp.K + (size_t)pid * bp.max_stages * 288;
    p.dU = bp.dU + (size_t)pid * bp.max_stages * 24;
    p.g0h0 = bp.g0h0 + (size_t)pid * 600;
    int rpos[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) rpos[q] = hkd::cr_dense_pos(min(lane + 32 * q, hkd::kCrNnz - 1));
    SolverState& st = bp.state[pid];
    SolveCtl& ctl = bp.ctl[pid];
    double reg = st.reg;
    __syncwarp();
    // MultiPhaseDDP::backward_sweep_regularized (Q8)
    bool success = false;
    int nsw = 0;
    double dV1 = 0.0, dV2 = 0.0;
    while (!success) {
        ++nsw;
        success = w1_backward_sweep(sm, p, reg, dt, rpos, dV1, dV2);
        if (success) break;
        reg = fmax(reg * opt.update_regularization, 1e-03);
        if (reg > 1e2) break;
    }
    reg = reg / 20;
    if (reg < 1e-06) reg = 0;
    if (lane == 0) {
        st.reg = reg; st.dV_1 = dV1; st.dV_2 = dV2; st.sweep_ok = success ? 1 : 0;
        ctl.n_sweeps += nsw;
        if (ctl.iter <= HSDDP_TRACE_CAP) {
            hsddp_iter_record& rec = bp.trace[(size_t)pid * HSDDP_TRACE_CAP + ctl.iter - 1];
            rec.n_sweeps = nsw; rec.reg_after = reg;
        }
        if (!success) { ctl.status = HSDDP_STATUS_REG_OVERFLOW; ctl.active = 0; }  // bad_solve (:321-324,421-427)
    }
}

}  // namespace hsddp
