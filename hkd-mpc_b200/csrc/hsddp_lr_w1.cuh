// Linear rollout with ONE WARP PER PROBLEM (phased driver): MultiPhaseDDP::linear_rollout + SinglePhase::linear_rollout
// (MultiPhaseDDP.cpp:20-50, SinglePhase.cpp:145-178) as its own kernel between the backward sweep and the forward phase.
//
// The dX recursion is sequential in time and one warp wide (lane roles as in linear_rollout_block, hsddp_sweep.cuh); inside
// the four-warp forward kernel it occupied a 34.8 KB block whose other three warps only fetched and waited, six problems per
// SM.  Here a problem costs one warp and a three-slot ring of stage data (12.8 KB): sixteen problems per SM walk their
// chains side by side, each fetching two stages ahead with cp.async (one commit group per stage).  The expected cost change
// needs nothing from HBM any more: the whole stage record sits in the ring slot, dx and du of the stage are in shared memory
// when it is evaluated, and the terminal terms are taken at the phase boundaries of the recursion.
// The recursion is bit-identical to linear_rollout_block; dV_1 / dV_2 are summed in a different order (rounding only).
#pragma once
#include "hsddp_sweep.cuh"

namespace hsddp {

constexpr int LW_REC = 288;               // slot: KT [24][12] | stage record (entries, lx, lu, luu: 196 doubles) | defect of node n+1 | dU
constexpr int LW_DF = LW_REC + 196;
constexpr int LW_DU = LW_DF + 24;
constexpr int LW_SLOT = LW_DU + 24;       // 532 doubles
constexpr int LW_RING = 3;
constexpr int LW_UNITS = LW_SLOT / 2;     // 16-byte units per slot: 144 | 98 | 12 | 12

struct LrW1 {
    alignas(16) double ring[LW_RING * LW_SLOT];
    double V[48];     // [dx (24) | coupled controls du_r (12) | 0 (9) | constants {0, dt, dt / m}]
    double duf[24];   // the stage's full control step (expected cost change)
    int n_phases, n_stages;
    int horizon[MAXPH], node_off[MAXPH], stage_off[MAXPH];
    unsigned cmask[MAXPH], nmask[MAXPH];
};

struct LrW1Ptrs {  // per-problem HBM pointers (registers, warp-uniform)
    const double *K, *lqg, *Defect, *dU, *tq;
    double *dX, *KdX, *U_t;
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// stage s (node n) into its ring slot; one commit group per call (also when there is nothing left to fetch).
// Four contiguous source runs (gains, record, defect, feed-forward); fixed trip counts, so every copy is one address add.
__device__ __forceinline__ void lw_prefetch(LrW1& sm, const LrW1Ptrs& p, int s, int n, bool valid) {
    if (valid) {
        const int lane = threadIdx.x & 31;
        double* slot = sm.ring + (s % LW_RING) * LW_SLOT + 2 * lane;
        const double* srcK = p.K + (size_t)s * 288 + 2 * lane;
        const double* srcR = p.lqg + (size_t)s * CR_STRIDE + 2 * lane;
#pragma unroll
        for (int it = 0; it < 4; ++it) cp_async16(slot + 64 * it, srcK + 64 * it);              // gains: units 0..127
        if (lane < 16) cp_async16(slot + 256, srcK + 256);                                     //        units 128..143
#pragma unroll
        for (int it = 0; it < 3; ++it) cp_async16(slot + LW_REC + 64 * it, srcR + 64 * it);    // record: units 0..95
        if (lane < 2) cp_async16(slot + LW_REC + 192, srcR + 192);                             //         units 96, 97
        if (lane < 12) {
            cp_async16(slot + LW_DF, p.Defect + 24 * (n + 1) + 2 * lane);
            cp_async16(slot + LW_DU, p.dU + 24 * s + 2 * lane);
        }
    }
    cp_async_commit();
}

// terminal terms of phase p at its end state V[0..23] (SinglePhase.cpp:174-177), component i = lane
__device__ __forceinline__ void lw_terminal(const LrW1& sm, const LrW1Ptrs& p, int ph, int i, double& dV1, double& dV2) {
    const unsigned cm = sm.cmask[ph];
    const double* trec = p.tq + ph * TQ_STRIDE;
    const double* dxv = sm.V;
    const double dxi = dxv[i];
    dV1 += trec[TQ_PHIX + i] * dxi;
    double qdx = weight_Qf(i, cm) * dxi;
    if (i >= 3 && i < 6) {
        for (int l = 0; l < 4; ++l) {
            const double c = (double)((cm >> l) & 1u);
            const double w = (20.0 * c * weight_foot(l, i - 3, cm)) * c;
            qdx += w * dxi - w * dxv[12 + 3 * l + i - 3];
        }
    } else if (i >= 12) {
        const int l = (i - 12) / 3, jj = (i - 12) % 3;
        const double c = (double)((cm >> l) & 1u);
        const double w = (20.0 * c * weight_foot(l, jj, cm)) * c;
        qdx += w * dxi - w * dxv[3 + jj];
    }
    for (int l = 0; l < 4; ++l) {
        const double wh = trec[TQ_WH + l];
        if (wh != 0.0) {
            double hd = 0.0;
            for (int j = 0; j < 24; ++j) hd = fma(trec[TQ_HX + 24 * l + j], dxv[j], hd);
            qdx += wh * trec[TQ_HX + 24 * l + i] * hd;
        }
    }
    dV2 += dxi * qdx;
}

// expected cost change of (stage, component i): lx dx + lu du and dx' lxx dx + du' luu du (SinglePhase.cpp:165-172); dx of the
// stage's node in V[0..23], the stage's control step in duf, the stage record in the ring slot.  The weights of the lane's
// row of lxx are per-phase constants (LwWeights, set at the phase start): wq on the diagonal and up to four foot-placement
// terms w (dx_i - dx_idx) -- four for a position row, one for a foot row, none (w = 0) elsewhere; same terms in the same order
// as lr_dv_elem (hsddp_sweep.cuh).
struct LwWeights {
    double wq, wr, wf[4];
    unsigned fidx;  // four byte indices into V
};
__device__ __forceinline__ LwWeights lw_weights(int i, unsigned cm, double dt) {
    LwWeights w;
    w.wq = (i < 24) ? dt * weight_Q(i, cm) : 0.0;
    w.wr = (i < 24) ? dt * weight_R(i) : 0.0;
    w.fidx = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l) w.wf[l] = 0.0;
    if (i >= 3 && i < 6) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const double c = (double)((cm >> l) & 1u);
            w.wf[l] = (dt * c * weight_foot(l, i - 3, cm)) * c;
            w.fidx |= (unsigned)(12 + 3 * l + i - 3) << (8 * l);
        }
    } else if (i >= 12 && i < 24) {
        const int l = (i - 12) / 3, jj = (i - 12) % 3;
        const double c = (double)((cm >> l) & 1u);
        w.wf[0] = (dt * c * weight_foot(l, jj, cm)) * c;
        w.fidx = (unsigned)(3 + jj);
    }
    return w;
}
__device__ __forceinline__ void lw_dv(const LrW1& sm, const double* rec, const LwWeights& w, int i, double& dV1, double& dV2) {
    const double* dxv = sm.V;
    const double* duv = sm.duf;
    const double dxi = dxv[i], dui = duv[i];
    dV1 += rec[CR_LX + i] * dxi + rec[CR_LU + i] * dui;
    double qdx = w.wq * dxi;
#pragma unroll
    for (int q = 0; q < 4; ++q) qdx += w.wf[q] * dxi - w.wf[q] * dxv[(w.fidx >> (8 * q)) & 255u];
    dV2 += dxi * qdx;
    double rdu = w.wr * dui;
    if (i < 12) {
        const int l = i / 3, a = i % 3;
#pragma unroll
        for (int b = 0; b < 3; ++b) rdu += rec[CR_LUU + 9 * l + 3 * a + b] * duv[3 * l + b];
    }
    dV2 += dui * rdu;
}

#ifndef HSDDP_LW_MINB
#define HSDDP_LW_MINB 16
#endif
__global__ void __launch_bounds__(32, HSDDP_LW_MINB) k_lr_w1(BatchPtrs bp, hsddp_options opt) {
    __shared__ LrW1 sm;
    const int lane = threadIdx.x;
    if (bp.n_active && (int)blockIdx.x >= *bp.n_active) return;  // (see BatchPtrs::active)
    const int pid = bp.active ? bp.active[blockIdx.x] : (int)blockIdx.x;
    if (!opt.MS || !bp.ctl[pid].active) return;  // (MultiPhaseDDP.cpp:326-329: only with multiple shooting; a failed sweep has ended the solve)
    const DevSchedule* sc = bp.sched + bp.sched_id[pid];
    if (lane == 0) { sm.n_phases = sc->n_phases; sm.n_stages = sc->n_stages; }
    if (lane < MAXPH) {
        sm.horizon[lane] = sc->horizon[lane]; sm.node_off[lane] = sc->node_off[lane]; sm.stage_off[lane] = sc->stage_off[lane];
        sm.cmask[lane] = sc->cmask[lane]; sm.nmask[lane] = sc->nmask[lane];
    }
    const double dt = sc->dt;
    const double eps = 1.0;
    LrW1Ptrs p;
    p.K = bp.K + (size_t)pid * bp.max_stages * 288;
    p.lqg = bp.lq + (size_t)pid * bp.max_stages * CR_STRIDE;
    p.Defect = bp.Defect + (size_t)pid * bp.max_nodes * 24;
    p.dU = bp.dU + (size_t)pid * bp.max_stages * 24;
    p.tq = bp.tq + (size_t)pid * MAXPH * TQ_STRIDE;
    p.dX = bp.dX + (size_t)pid * bp.max_nodes * 24;
    p.KdX = bp.KdX + (size_t)pid * bp.max_stages * 12;
    p.U_t = bp.U_t + (size_t)pid * bp.max_stages * 24;
    double* V = sm.V;
    double* lrc = V + 45;  // constants {0, dt, dt / m} of the sparse rows
    if (lane < 9) V[36 + lane] = 0.0;
    if (lane == 0) { lrc[0] = 0.0; lrc[1] = dt; lrc[2] = (1.0 / hkd::kMass) * dt; }
    __syncwarp();
    const int N = sm.n_stages, NP = sm.n_phases;
    // the ring runs LW_RING - 1 stages ahead; `pph` is the phase of the stage being fetched
    int pph = 0;
    auto node_of = [&](int s) {
        while (pph + 1 < NP && s >= sm.stage_off[pph + 1]) ++pph;
        return sm.node_off[pph] + s - sm.stage_off[pph];
    };
#pragma unroll
    for (int s = 0; s < LW_RING - 1; ++s) lw_prefetch(sm, p, s, s < N ? node_of(s) : 0, s < N);
    int ph = -1;
    unsigned cm = 0;
    unsigned long long cpack = 0, vpack = 0;
    LwWeights wts = lw_weights(24, 0u, dt);
    double dx = 0.0, dV1 = 0.0, dV2 = 0.0;
    for (int s = 0; s < N; ++s) {
        {
            const int sp = s + LW_RING - 1;
            lw_prefetch(sm, p, sp, sp < N ? node_of(sp) : 0, sp < N);
        }
        if (ph < 0 || (ph + 1 < NP && s >= sm.stage_off[ph + 1])) {
            // ---- phase start ----
            if (ph >= 0 && lane < 24) lw_terminal(sm, p, ph, lane, dV1, dV2);  // terminal terms of the phase that ends here
            ++ph;
            while (ph + 1 < NP && s >= sm.stage_off[ph + 1]) ++ph;
            cm = sm.cmask[ph];
            wts = lw_weights(lane, cm, dt);
            const int n = sm.node_off[ph];
            // dx_init = Px dX_end(prev) (zero for the first phase); dX[0] = dx_init + eps Defect[0]
            double dxi = 0.0;
            if (ph > 0 && lane < 24) {
                const unsigned pc_ = sm.cmask[ph - 1], pn_ = sm.nmask[ph - 1];
                dxi = dx;
                if (lane >= 12) {
                    const int l = (lane - 12) / 3, r = (lane - 12) % 3;
                    const bool cl = (pc_ >> l) & 1u, nl = (pn_ >> l) & 1u;
                    if (cl && !nl) dxi = 0.0;
                    if (!cl && nl) {
                        if (r == 2) dxi = 0.0;
                        else {
                            const double* Jc = p.tq + (ph - 1) * TQ_STRIDE + TQ_JC + 18 * l + 6 * r;
                            double acc = V[3 + r];
#pragma unroll
                            for (int c = 0; c < 3; ++c) acc = fma(Jc[c], V[c], acc);
#pragma unroll
                            for (int c = 0; c < 3; ++c) acc = fma(Jc[3 + c], V[12 + 3 * l + c], acc);
                            dxi = acc;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane < 24) {
                dx = dxi + eps * p.Defect[24 * n + lane];
                p.dX[24 * n + lane] = dx;
                V[lane] = dx;
            }
            // per-lane description of the sparse rows of [A - I | B_r] that are NOT the angular-acceleration rows
            // (linear_rollout_block, hsddp_sweep.cuh): dx+_i = dx_i + sum_q coef_q V[vidx_q]
            {
                unsigned char ci[5] = {100, 100, 100, 100, 100}, vi[5] = {36, 36, 36, 36, 36};
                if (lane == 0) { ci[0] = 0; ci[1] = 1; ci[2] = 2; ci[3] = 3; vi[0] = 1; vi[1] = 2; vi[2] = 7; vi[3] = 8; }
                else if (lane == 1) { ci[0] = 4; ci[1] = 5; ci[2] = 6; vi[0] = 2; vi[1] = 7; vi[2] = 8; }
                else if (lane == 2) { ci[0] = 7; ci[1] = 8; ci[2] = 9; ci[3] = 10; ci[4] = 11; vi[0] = 1; vi[1] = 2; vi[2] = 6; vi[3] = 7; vi[4] = 8; }
                else if (lane < 6) { ci[0] = 101; vi[0] = (unsigned char)(lane + 6); }
                else if (lane >= 9 && lane < 12) {
                    for (int l = 0; l < 4; ++l) { ci[l] = ((cm >> l) & 1u) ? 102 : 100; vi[l] = (unsigned char)(24 + 3 * l + lane - 9); }
                } else if (lane >= 12 && lane < 24) {
                    if (!((cm >> ((lane - 12) / 3)) & 1u)) { ci[0] = 101; vi[0] = (unsigned char)(24 + lane - 12); }
                }
                cpack = 0; vpack = 0;
                for (int q = 0; q < 5; ++q) { cpack |= (unsigned long long)ci[q] << (8 * q); vpack |= (unsigned long long)vi[q] << (8 * q); }
            }
            __syncwarp();
        }
        cp_async_wait_group<LW_RING - 1>();
        __syncwarp();
        const int k = s - sm.stage_off[ph];
        const int n = sm.node_off[ph] + k;
        const double* slot = sm.ring + (s % LW_RING) * LW_SLOT;
        const double* KT = slot;
        const double* Rc = slot + LW_REC;  // the stage record: compact entries of [A - I | B_r], then lx, lu, luu
        const double* dfn = slot + LW_DF;
        const double* dUs = slot + LW_DU;
        // ---- phase A: feedback K_r dx (lanes 0..23) and the state part of the angular-acceleration rows (lanes 24..29) ----
        double pa = 0.0, pb = 0.0;
        if (lane < 24) {
            const int c = (lane < 12) ? lane : lane - 12, j0 = (lane < 12) ? 0 : 12;
            const double* kt = KT + j0 * 12 + c;
            const double* v = V + j0;
#pragma unroll
            for (int j = 0; j < 12; j += 2) { pa = fma(kt[j * 12], v[j], pa); pb = fma(kt[(j + 1) * 12], v[j + 1], pb); }
        } else if (lane < 30) {
            const int a = (lane - 24) >> 1;
            const double* W = Rc + 12 + 29 * a;
            if ((lane & 1) == 0) {
#pragma unroll
                for (int q = 0; q < 8; q += 2) { pa = fma(W[q], V[q], pa); pb = fma(W[q + 1], V[q + 1], pb); }
                pa = fma(W[8], V[8], pa);
            } else {
#pragma unroll
                for (int l = 0; l < 4; ++l) { pa = fma(W[9 + 2 * l], V[12 + 3 * l], pa); pb = fma(W[10 + 2 * l], V[13 + 3 * l], pb); }
            }
        }
        const double part = pa + pb;
        const double hi = __shfl_down_sync(0xffffffffu, part, 12);
        const int l6 = 24 + 2 * min(max(lane - 6, 0), 2);
        const double w0 = __shfl_sync(0xffffffffu, part, l6), w1 = __shfl_sync(0xffffffffu, part, l6 + 1);
        if (lane < 12) {
            const int i = act_index(lane, cm);
            const double kdx = part + hi;
            const double du = eps * dUs[i] + kdx;
            p.KdX[12 * s + lane] = kdx;  // kept for the trial rollouts of this iteration (hybrid_rollout_block<true>)
            V[24 + lane] = du;
            sm.duf[i] = du;
            p.U_t[24 * s + i] = du;
        } else if (lane < 24) {
            const int i = inact_index(lane - 12, cm);
            const double du = eps * dUs[i];
            sm.duf[i] = du;
            p.U_t[24 * s + i] = du;
        }
        __syncwarp();
        // ---- expected cost change of the stage (off the recursion's chain) ----
        if (lane < 24) lw_dv(sm, Rc, wts, lane, dV1, dV2);
        // ---- phase B: dx+ = dx + (A - I) dx + B_r du_r + eps d ----
        if (lane < 24) {
            double acc = dx;
            if (lane >= 6 && lane < 9) {  // angular acceleration: the state part from phase A, then the 12 coupled controls
                const double* W = Rc + 12 + 29 * (lane - 6) + 17;
                double b0 = 0.0, b1 = 0.0;
#pragma unroll
                for (int c = 0; c < 12; c += 2) { b0 = fma(W[c], V[24 + c], b0); b1 = fma(W[c + 1], V[25 + c], b1); }
                acc += (w0 + w1) + (b0 + b1);
            } else {
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const unsigned ci = (unsigned)(cpack >> (8 * q)) & 255u, vi = (unsigned)(vpack >> (8 * q)) & 255u;
                    const double coef = (ci < 100u) ? Rc[ci] : lrc[ci - 100u];
                    acc = fma(coef, V[vi], acc);
                }
            }
            dx = acc + eps * dfn[lane];
        }
        __syncwarp();
        if (lane < 24) { p.dX[24 * (n + 1) + lane] = dx; V[lane] = dx; }
        __syncwarp();
    }
    if (ph >= 0 && lane < 24) lw_terminal(sm, p, ph, lane, dV1, dV2);  // the last phase
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { dV1 += __shfl_xor_sync(0xffffffffu, dV1, o); dV2 += __shfl_xor_sync(0xffffffffu, dV2, o); }
    if (lane == 0) { bp.state[pid].dV_1 = dV1; bp.state[pid].dV_2 = dV2; }
}

}  // namespace hsddp
