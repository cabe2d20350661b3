// HKD (hybrid kino-dynamics) Mini Cheetah model as sm_100a __device__ functions.
//
// Replaces the CasADi-generated C of the reference
//   hkinodyn               HKDMPC/HKD-TrajOpt/CasadiGen/source/hkinodyn_casadi.cpp:177-658
//   hkinodyn_par           .../hkinodyn_par_casadi.cpp:181-2800
//   compute_foot_position  .../comp_foot_pos_casadi.cpp:45-160
//   comp_foot_jacob_{1..4} .../comp_foot_jacob_k_casadi.cpp:45-520
// and the marshalling around it (common/casadi_interface.cpp:5-79, HKDModel.h:33-61).
// The functions are re-derived from the model's physics (single rigid body with
// ZYX Euler angles, world-frame ground reaction forces on kinematic legs, explicit
// Euler step) with hand-derived ANALYTIC Jacobians that expose the structure the
// Riccati kernel exploits:
//     A = I + At,  At non-zero only in rows {0,1,2} (Euler rates), {3,4,5} (dt at
//         columns 9..11) and {6,7,8} (angular acceleration)
//     B : rows {6,7,8} x cols 0..11 dense; (9+j, 3l+j) = c_l dt/m; (12+i,12+i) = (1-c_l) dt
// The inertia products of order 1e-19 that the generated code carries (numerical
// residue of a matrix inverse) are dropped; their effect is below one ulp.
// Everything is __host__ __device__ so the host-side problem assembly
// (compute_hkd_state) and the unit tests run the very same code.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define HKD_HD __host__ __device__ __forceinline__
#else
#define HKD_HD inline
#endif

namespace hkd {

constexpr double kMass = 8.9120000000000008e+00;
constexpr double kGravZ = -9.8100000000000005e+00;
constexpr double kIxx = 2.7460779999999994e-02, kIyy = 2.4251579680000002e-01, kIzz = 2.6519357680000000e-01;
constexpr double kJxx = 3.6415571589736352e+01, kJyy = 4.1234427331951844e+00, kJzz = 3.7708303951651367e+00;
constexpr double kHipX = 1.9000000000000000e-01, kHipY = 4.9000000000000002e-02;
constexpr double kAbad = 6.2000000000000000e-02, kThigh = -2.0899999999999999e-01, kShank = -1.9500000000000001e-01;

// Linearisation record of one stage.  In HBM it is COMPACT: only the 111 structurally non-zero, state-dependent
// entries of rows 0..11 of [A - I | B_r] are stored (B_r = the 24x12 matrix of the COUPLED controls, reduced column
// c = 3*leg+j; only stance-leg columns have entries in these rows).  Order of the compact record Rc[112]:
//     [ 0.. 3]  row 0 (yaw rate)    columns {1,2,7,8}
//     [ 4.. 6]  row 1 (pitch rate)  columns {2,7,8}
//     [ 7..11]  row 2 (roll rate)   columns {1,2,6,7,8}
//     [12+29a ..] rows 6+a (angular acceleration), a = 0..2:  +0..8 columns 0..8 ; +9+2l+k column 12+3l+k (foot x,y of leg l) ;
//                 +17+c column 24+c of the dense tile = B_r column c
//     [99+4j+l]  row 9+j, B_r column 3l+j  ((c_l / m) dt)
//     [111]      padding
// The rows 3..5 of A - I hold the constant dt at columns 9..11 and are not stored.
// In shared memory the Riccati kernel works on the dense tile R[12][44] (row stride 44 = 12 mod 16 doubles:
// conflict-free tensor-core fragment loads; columns 0..23 = A - I, 24..35 = B_r, 36..43 zero padding); the compact
// entries are scattered into it by cp.async (cr_dense_pos), everything else in the tile stays zero.
constexpr int kRld = 44;
constexpr int kRSize = 12 * kRld;
constexpr int kCrNnz = 111;
constexpr int kCrSize = 112;
HKD_HD int cr_w(int a, int q) { return 12 + 29 * a + q; }            // row 6+a: q = column (0..8), 9+2l+k (foot), 17+c (B_r)
HKD_HD int cr_v(int j, int l) { return 99 + 4 * j + l; }             // row 9+j, B_r column 3l+j
// position of compact entry i inside the dense tile R[12][kRld]
HKD_HD int cr_dense_pos(int i) {
    if (i < 4) return (i < 2) ? 1 + i : 5 + i;                        // row 0: 1,2,7,8
    if (i < 7) return kRld + ((i == 4) ? 2 : 3 + i - 1);              // row 1: 2,7,8   (i=5 -> 7, i=6 -> 8)
    if (i < 12) return 2 * kRld + ((i < 9) ? i - 6 : i - 3);          // row 2: 1,2,6,7,8  (i=7->1, 8->2, 9->6, 10->7, 11->8)
    if (i < 99) {
        const int a = (i - 12) / 29, q = (i - 12) % 29;
        const int col = (q < 9) ? q : (q < 17) ? 12 + 3 * ((q - 9) / 2) + (q - 9) % 2 : 24 + (q - 17);
        return (6 + a) * kRld + col;
    }
    const int j = (i - 99) / 4, l = (i - 99) % 4;
    return (9 + j) * kRld + 24 + 3 * l + j;
}

struct Trig {
    double sy, cy, sp, cp, sr, cr;
};

#ifdef __CUDACC__
// One out-of-line copy of the FP64 sincos / log expansions (argument reduction + slow path are ~200
// instructions each): the solver kernel is instruction-cache bound when they are inlined at every call site.
struct SinCos { double s, c; };
static __device__ __noinline__ SinCos sincos_nl(double a) {
    SinCos r;
    sincos(a, &r.s, &r.c);
    return r;
}
static __device__ __noinline__ double log_nl(double a) { return log(a); }
#endif

HKD_HD Trig trig_of(double yaw, double pitch, double roll) {
    Trig t;
#ifdef __CUDA_ARCH__
    const SinCos a = sincos_nl(yaw), b = sincos_nl(pitch), c = sincos_nl(roll);
    t.sy = a.s; t.cy = a.c; t.sp = b.s; t.cp = b.c; t.sr = c.s; t.cr = c.c;
#else
    t.sy = sin(yaw); t.cy = cos(yaw); t.sp = sin(pitch); t.cp = cos(pitch); t.sr = sin(roll); t.cr = cos(roll);
#endif
    return t;
}

// R = Rz(yaw) Ry(pitch) Rx(roll), row-major R[3*i+j]
HKD_HD void rotation(const Trig& t, double R[9]) {
    R[0] = t.cy * t.cp; R[1] = t.cy * t.sp * t.sr - t.sy * t.cr; R[2] = t.sy * t.sr + t.cy * t.sp * t.cr;
    R[3] = t.sy * t.cp; R[4] = t.cy * t.cr + t.sy * t.sp * t.sr; R[5] = t.sy * t.sp * t.cr - t.cy * t.sr;
    R[6] = -t.sp;       R[7] = t.cp * t.sr;                      R[8] = t.cp * t.cr;
}

// world-frame net force and net torque about the CoM of the stance legs
HKD_HD void wrench(const double* x, const double* u, unsigned cmask, double F[3], double tau[3]) {
    F[0] = F[1] = F[2] = 0.0;
    tau[0] = tau[1] = tau[2] = 0.0;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        if ((cmask >> l) & 1u) {
            const double rx = x[12 + 3 * l] - x[3], ry = x[13 + 3 * l] - x[4], rz = -x[5];  // stance foot on z = 0
            const double fx = u[3 * l], fy = u[3 * l + 1], fz = u[3 * l + 2];
            tau[0] += ry * fz - rz * fy;
            tau[1] += rz * fx - rx * fz;
            tau[2] += rx * fy - ry * fx;
            F[0] += fx; F[1] += fy; F[2] += fz;
        }
    }
}

// xn = x + dt f(x,u)  (HKD::Model::dynamics)
HKD_HD void dynamics(const double* x, const double* u, double dt, unsigned cmask, double* xn) {
    const Trig t = trig_of(x[0], x[1], x[2]);
    const double wx = x[6], wy = x[7], wz = x[8];
    const double s1 = t.sr * wy + t.cr * wz;
    const double s2 = t.cr * wy - t.sr * wz;
    const double icp = 1.0 / t.cp;
    xn[0] = x[0] + (s1 * icp) * dt;
    xn[1] = x[1] + s2 * dt;
    xn[2] = x[2] + (wx + (t.sp * icp) * s1) * dt;
    xn[3] = x[3] + x[9] * dt;
    xn[4] = x[4] + x[10] * dt;
    xn[5] = x[5] + x[11] * dt;
    double R[9], F[3], tw[3];
    rotation(t, R);
    wrench(x, u, cmask, F, tw);
    // body-frame torque: gyroscopic term + R^T tau_world
    const double tbx = (kIyy - kIzz) * wy * wz + (R[0] * tw[0] + R[3] * tw[1] + R[6] * tw[2]);
    const double tby = (kIzz - kIxx) * wz * wx + (R[1] * tw[0] + R[4] * tw[1] + R[7] * tw[2]);
    const double tbz = (kIxx - kIyy) * wx * wy + (R[2] * tw[0] + R[5] * tw[1] + R[8] * tw[2]);
    xn[6] = wx + (kJxx * tbx) * dt;
    xn[7] = wy + (kJyy * tby) * dt;
    xn[8] = wz + (kJzz * tbz) * dt;
    xn[9] = x[9] + (F[0] / kMass) * dt;
    xn[10] = x[10] + (F[1] / kMass) * dt;
    xn[11] = x[11] + (kGravZ + F[2] / kMass) * dt;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const double sw = ((cmask >> l) & 1u) ? 0.0 : 1.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) xn[12 + 3 * l + j] = x[12 + 3 * l + j] + (sw * u[12 + 3 * l + j]) * dt;
    }
}

// Analytic linearisation (HKD::Model::dynamics_partial) written straight into a COMPACT stage record Rc[kCrSize] (layout above).
// PARTS bit 0: the Euler-rate rows and the body columns (0..8) of the angular-acceleration rows;
// PARTS bit 1: the per-leg columns (foot x,y and B_r) of the angular-acceleration rows and the linear-acceleration rows.
// The two halves write disjoint entries, so two threads can share one stage (lq_approximation_block).
template <int PARTS>
HKD_HD void dynamics_partial_parts(const double* x, const double* u, double dt, unsigned cmask, double* Rc) {
    const Trig t = trig_of(x[0], x[1], x[2]);
    double R[9];
    rotation(t, R);
    const double jd[3] = {dt * kJxx, dt * kJyy, dt * kJzz};
    double M[9];  // M = dt * Jinv * R^T
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i) M[3 * a + i] = jd[a] * R[3 * i + a];
    if (PARTS & 1) {
        const double wx = x[6], wy = x[7], wz = x[8];
        const double s1 = t.sr * wy + t.cr * wz;
        const double s2 = t.cr * wy - t.sr * wz;
        const double icp = 1.0 / t.cp;
        const double tp = t.sp * icp;
        // Euler-rate rows 0..2
        Rc[0] = dt * (s1 * t.sp * icp * icp);
        Rc[1] = dt * (s2 * icp);
        Rc[2] = dt * (t.sr * icp);
        Rc[3] = dt * (t.cr * icp);
        Rc[4] = dt * (-s1);
        Rc[5] = dt * t.cr;
        Rc[6] = dt * (-t.sr);
        Rc[7] = dt * (s1 * icp * icp);
        Rc[8] = dt * (tp * s2);
        Rc[9] = dt;
        Rc[10] = dt * (tp * t.sr);
        Rc[11] = dt * (tp * t.cr);
        // position rows 3..5 hold the constant dt at columns 9..11: not stored
        // angular-acceleration rows 6..8, body columns
        double F[3], tw[3];
        wrench(x, u, cmask, F, tw);
        // d/d yaw:  (dR/dyaw)^T tau = -R[1][:] tau0 + R[0][:] tau1 ; d/d roll: (0, (R^T tau)_2, -(R^T tau)_1)
        const double rt1 = R[1] * tw[0] + R[4] * tw[1] + R[7] * tw[2];
        const double rt2 = R[2] * tw[0] + R[5] * tw[1] + R[8] * tw[2];
        const double dP[9] = {-t.cy * t.sp, t.cy * t.cp * t.sr, t.cy * t.cp * t.cr,
                              -t.sy * t.sp, t.sy * t.cp * t.sr, t.sy * t.cp * t.cr,
                              -t.cp,        -t.sp * t.sr,       -t.sp * t.cr};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double* row = Rc + cr_w(a, 0);
            row[0] = jd[a] * (-R[3 + a] * tw[0] + R[a] * tw[1]);
            row[1] = jd[a] * (dP[a] * tw[0] + dP[3 + a] * tw[1] + dP[6 + a] * tw[2]);
            // position columns: tau_world depends on p through r_l = foot - p  ->  column j = M (F x e_j)
            row[3] = M[3 * a + 1] * F[2] - M[3 * a + 2] * F[1];
            row[4] = -M[3 * a + 0] * F[2] + M[3 * a + 2] * F[0];
            row[5] = M[3 * a + 0] * F[1] - M[3 * a + 1] * F[0];
        }
        Rc[cr_w(0, 2)] = 0.0;
        Rc[cr_w(1, 2)] = jd[1] * rt2;
        Rc[cr_w(2, 2)] = jd[2] * (-rt1);
        // gyroscopic block
        Rc[cr_w(0, 6)] = 0.0;
        Rc[cr_w(0, 7)] = jd[0] * (kIyy - kIzz) * wz;
        Rc[cr_w(0, 8)] = jd[0] * (kIyy - kIzz) * wy;
        Rc[cr_w(1, 6)] = jd[1] * (kIzz - kIxx) * wz;
        Rc[cr_w(1, 7)] = 0.0;
        Rc[cr_w(1, 8)] = jd[1] * (kIzz - kIxx) * wx;
        Rc[cr_w(2, 6)] = jd[2] * (kIxx - kIyy) * wy;
        Rc[cr_w(2, 7)] = jd[2] * (kIxx - kIyy) * wx;
        Rc[cr_w(2, 8)] = 0.0;
    }
    if (PARTS & 2) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const bool stance = (cmask >> l) & 1u;
            const double rx = x[12 + 3 * l] - x[3], ry = x[13 + 3 * l] - x[4], rz = -x[5];
            const double fx = u[3 * l], fy = u[3 * l + 1], fz = u[3 * l + 2];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double m0 = M[3 * a], m1 = M[3 * a + 1], m2 = M[3 * a + 2];
                // foot x,y columns: M (e_j x f) ; force columns: M (r x e_j)   (zero for a swing leg)
                Rc[cr_w(a, 9 + 2 * l)] = stance ? (-m1 * fz + m2 * fy) : 0.0;
                Rc[cr_w(a, 10 + 2 * l)] = stance ? (m0 * fz - m2 * fx) : 0.0;
                Rc[cr_w(a, 17 + 3 * l + 0)] = stance ? (m1 * rz - m2 * ry) : 0.0;
                Rc[cr_w(a, 17 + 3 * l + 1)] = stance ? (-m0 * rz + m2 * rx) : 0.0;
                Rc[cr_w(a, 17 + 3 * l + 2)] = stance ? (m0 * ry - m1 * rx) : 0.0;
            }
            // linear acceleration rows 9..11: (c/m) dt on the leg's own force component
#pragma unroll
            for (int j = 0; j < 3; ++j) Rc[cr_v(j, l)] = stance ? (1.0 / kMass) * dt : 0.0;
        }
    }
}
HKD_HD void dynamics_partial_record(const double* x, const double* u, double dt, unsigned cmask, double* Rc) {
    dynamics_partial_parts<3>(x, u, dt, cmask, Rc);
}

// expand a compact stage record to the dense column-major A, B of the reference
HKD_HD void expand_AB(const double* Rc, double dt, unsigned cmask, double* A, double* B) {
    for (int i = 0; i < 576; ++i) { A[i] = 0.0; B[i] = 0.0; }
    for (int i = 0; i < 24; ++i) A[i * 25] = 1.0;
    for (int j = 0; j < 3; ++j) A[(3 + j) + 24 * (9 + j)] += dt;
    for (int i = 0; i < kCrNnz; ++i) {
        const int pos = cr_dense_pos(i), r = pos / kRld, c = pos % kRld;
        if (c < 24) A[r + 24 * c] += Rc[i];
        else if ((cmask >> ((c - 24) / 3)) & 1u) B[r + 24 * (c - 24)] = Rc[i];
    }
    for (int l = 0; l < 4; ++l)
        if (!((cmask >> l) & 1u))
            for (int j = 0; j < 3; ++j) B[(12 + 3 * l + j) + 24 * (12 + 3 * l + j)] = dt;
}

// leg kinematics in the body frame and its joint Jacobian dq[3*i+j] = d pb_i / d q_j
HKD_HD void leg_kinematics(const double* q, int leg, double pb[3], double* dq) {
    const double side = (leg & 1) ? 1.0 : -1.0;
    const double fore = (leg < 2) ? 1.0 : -1.0;
    const double l1 = kAbad * side;
    double s1, c1, s2, c2, s3, c3;
#ifdef __CUDA_ARCH__
    { const SinCos a = sincos_nl(q[0]), b = sincos_nl(-q[1]), c = sincos_nl(-q[2]); s1 = a.s; c1 = a.c; s2 = b.s; c2 = b.c; s3 = c.s; c3 = c.c; }
#else
    s1 = sin(q[0]); c1 = cos(q[0]); s2 = sin(-q[1]); c2 = cos(-q[1]); s3 = sin(-q[2]); c3 = cos(-q[2]);
#endif
    const double s23 = c2 * s3 + s2 * c3;
    const double c23 = c2 * c3 - s2 * s3;
    const double rho = kThigh * c2 + kShank * c23;
    const double sig = kThigh * s2 + kShank * s23;
    pb[0] = kHipX * fore + sig;
    pb[1] = kHipY * side + (c1 * l1 - s1 * rho);
    pb[2] = c1 * rho + s1 * l1;
    if (dq) {
        dq[0] = 0.0;                   dq[1] = -rho;        dq[2] = -(kShank * c23);
        dq[3] = -s1 * l1 - c1 * rho;   dq[4] = -(s1 * sig); dq[5] = -(s1 * (kShank * s23));
        dq[6] = c1 * l1 - s1 * rho;    dq[7] = c1 * sig;    dq[8] = c1 * (kShank * s23);
    }
}

HKD_HD void foot_position(const double* pos, const double* eul, const double* q, int leg, double p[3]) {
    double pb[3], R[9];
    leg_kinematics(q, leg, pb, nullptr);
    rotation(trig_of(eul[0], eul[1], eul[2]), R);
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = R[3 * i] * pb[0] + R[3 * i + 1] * pb[1] + R[3 * i + 2] * pb[2] + pos[i];
}

// Foot Jacobian restricted to its 9 structurally non-zero columns:
// Jc[3*i + c], c = 0..2 d/d eul(yaw,pitch,roll), 3..5 d/d qleg.  d/d pos = I.
HKD_HD void foot_jacobian_compact(const double* eul, const double* q, int leg, double Jc[18]) {
    double pb[3], dq[9], R[9];
    leg_kinematics(q, leg, pb, dq);
    const Trig t = trig_of(eul[0], eul[1], eul[2]);
    rotation(t, R);
    const double dP[9] = {-t.cy * t.sp, t.cy * t.cp * t.sr, t.cy * t.cp * t.cr,
                          -t.sy * t.sp, t.sy * t.cp * t.sr, t.sy * t.cp * t.cr,
                          -t.cp,        -t.sp * t.sr,       -t.sp * t.cr};
    // yaw: dRow0 = -Row1, dRow1 = Row0, dRow2 = 0
    Jc[0 * 6 + 0] = -(R[3] * pb[0] + R[4] * pb[1] + R[5] * pb[2]);
    Jc[1 * 6 + 0] = R[0] * pb[0] + R[1] * pb[1] + R[2] * pb[2];
    Jc[2 * 6 + 0] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        Jc[i * 6 + 1] = dP[3 * i] * pb[0] + dP[3 * i + 1] * pb[1] + dP[3 * i + 2] * pb[2];
        // roll: dCol1 = Col2, dCol2 = -Col1
        Jc[i * 6 + 2] = R[3 * i + 2] * pb[1] - R[3 * i + 1] * pb[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) Jc[i * 6 + 3 + j] = R[3 * i] * dq[j] + R[3 * i + 1] * dq[3 + j] + R[3 * i + 2] * dq[6 + j];
    }
}

// compute_hkd_state (HKDModel.h:65-96)
HKD_HD void hkd_state(const double* eul, const double* pos, const double* qJ, unsigned cmask, double* qdummy) {
    for (int l = 0; l < 4; ++l) {
        if ((cmask >> l) & 1u) foot_position(pos, eul, qJ + 3 * l, l, qdummy + 3 * l);
        else for (int j = 0; j < 3; ++j) qdummy[3 * l + j] = qJ[3 * l + j];
    }
}

}  // namespace hkd
