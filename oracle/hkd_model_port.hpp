// ORACLE (test infrastructure only — never linked by the product).
//
// Independent plain-C++ restatement of the Mini Cheetah hybrid kino-dynamics
// (HKD) model that the reference ships as CasADi-generated C:
//   hkinodyn               HKDMPC/HKD-TrajOpt/CasadiGen/source/hkinodyn_casadi.cpp:177-658
//   hkinodyn_par           .../hkinodyn_par_casadi.cpp:181-2800      (A 24x24 dense, B 60 nnz)
//   compute_foot_position  .../comp_foot_pos_casadi.cpp:45-160
//   comp_foot_jacob_{1..4} .../comp_foot_jacob_k_casadi.cpp:45-520   (3x18, cols [pos eul qJ(12)])
// It is written from the physics those expressions encode (single rigid body
// with ZYX Euler angles + kinematic legs, explicit Euler step), not from the
// generated statement list.  Jacobians come from forward-mode dual numbers, so
// this port is independent of the hand-derived analytic Jacobians used by the
// sm_100a kernels.  Pinning: tests/test_oracle_model.py checks it against the
// reference's own compiled CasADi code (oracle/_ref) and against golden vectors
// generated from that code (tests/golden/model_vectors.npz).
#pragma once
#include <cmath>

namespace hkd_port {

// ---- physical parameters (the numeric constants of the generated code) ----
constexpr double kMass = 8.9120000000000008e+00;                // hkinodyn_casadi.cpp (a68)
constexpr double kGrav = -9.8100000000000005e+00;
// body inertia (upper triangle; Iyz is structurally absent)
constexpr double kIxx = 2.7460779999999994e-02, kIyy = 2.4251579680000002e-01, kIzz = 2.6519357680000000e-01;
constexpr double kIxy = 1.0842021724855044e-19, kIxz = -1.2037062152420224e-35;
// inverse inertia
constexpr double kJxx = 3.6415571589736352e+01, kJyy = 4.1234427331951844e+00, kJzz = 3.7708303951651367e+00;
constexpr double kJxy = -1.6280111378663628e-17, kJxz = 1.6528925920107902e-33, kJyz = -7.3894969432494111e-52;
// leg geometry (comp_foot_pos_casadi.cpp)
constexpr double kHipX = 1.9000000000000000e-01, kHipY = 4.9000000000000002e-02;
constexpr double kAbad = 6.2000000000000000e-02, kThigh = -2.0899999999999999e-01, kShank = -1.9500000000000001e-01;

// ---- forward-mode dual number with N directions ----
template <int N>
struct Dual {
    double v;
    double d[N];
    Dual() : v(0) { for (int i = 0; i < N; ++i) d[i] = 0; }
    Dual(double c) : v(c) { for (int i = 0; i < N; ++i) d[i] = 0; }
};
template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v / b.v; for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v; return r; }
template <int N> inline Dual<N> sin(const Dual<N>& a) { Dual<N> r; r.v = std::sin(a.v); double c = std::cos(a.v); for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i]; return r; }
template <int N> inline Dual<N> cos(const Dual<N>& a) { Dual<N> r; r.v = std::cos(a.v); double s = -std::sin(a.v); for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }
template <int N> inline Dual<N> operator*(double a, const Dual<N>& b) { return Dual<N>(a) * b; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double b) { return a * Dual<N>(b); }
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { return a + Dual<N>(b); }
template <int N> inline Dual<N> operator+(double a, const Dual<N>& b) { return Dual<N>(a) + b; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { return a - Dual<N>(b); }
template <int N> inline Dual<N> operator-(double a, const Dual<N>& b) { return Dual<N>(a) - b; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) { return a / Dual<N>(b); }

using std::sin;
using std::cos;

// Rotation world<-body for ZYX Euler angles eul = (yaw, pitch, roll).
template <class S>
inline void rotation_zyx(const S& yaw, const S& pitch, const S& roll, S R[3][3]) {
    S cy = cos(yaw), sy = sin(yaw), cp = cos(pitch), sp = sin(pitch), cr = cos(roll), sr = sin(roll);
    R[0][0] = cy * cp; R[0][1] = cy * sp * sr - sy * cr; R[0][2] = sy * sr + cy * sp * cr;
    R[1][0] = sy * cp; R[1][1] = cy * cr + sy * sp * sr; R[1][2] = sy * sp * cr - cy * sr;
    R[2][0] = -sp;     R[2][1] = cp * sr;                R[2][2] = cp * cr;
}

// One explicit-Euler step of the HKD model.
//   x = [eul(yaw,pitch,roll) pos omega_body v_world qdummy(12)], u = [GRF(12) qJd(12)]
//   c[l] in {0,1}: stance flag of leg l (passed as double like HKDModel.h:39-41).
template <class S>
inline void hkd_step(const S x[24], const S u[24], double dt, const double c[4], S xn[24]) {
    const S &yaw = x[0], &pitch = x[1], &roll = x[2];
    const S &wx = x[6], &wy = x[7], &wz = x[8];
    S cp = cos(pitch), sp = sin(pitch), cr = cos(roll), sr = sin(roll);
    // Euler-angle rates from body angular velocity
    S yaw_d = (sr / cp) * wy + (cr / cp) * wz;
    S pitch_d = cr * wy - sr * wz;
    S roll_d = wx + (sr * sp / cp) * wy + (cr * sp / cp) * wz;
    xn[0] = yaw + yaw_d * dt;
    xn[1] = pitch + pitch_d * dt;
    xn[2] = roll + roll_d * dt;
    for (int i = 0; i < 3; ++i) xn[3 + i] = x[3 + i] + x[9 + i] * dt;

    S R[3][3];
    rotation_zyx(yaw, pitch, roll, R);
    // gyroscopic term  -(w x I w)
    S tx = (kIxy * wz - kIxz * wy) * wx + kIyy * wz * wy - kIzz * wy * wz;
    S ty = (kIxz * wx - kIxx * wz) * wx - (kIxy * wz) * wy + (kIzz * wx - kIxz * wz) * wz;
    S tz = (kIxx * wy - kIxy * wx) * wx + (kIxy * wy - kIyy * wx) * wy + kIxz * wy * wz;
    S fsum[3] = {S(0.0), S(0.0), S(0.0)};
    S gx(0.0), gy(0.0), gz(0.0);
    for (int l = 0; l < 4; ++l) {
        // lever arm in world frame; the stance foot is taken on the ground plane z = 0
        S rw[3] = {x[12 + 3 * l] - x[3], x[13 + 3 * l] - x[4], S(0.0) - x[5]};
        S fw[3] = {u[3 * l], u[3 * l + 1], u[3 * l + 2]};
        S rb[3], fb[3];
        for (int j = 0; j < 3; ++j) {
            rb[j] = R[0][j] * rw[0] + R[1][j] * rw[1] + R[2][j] * rw[2];
            fb[j] = R[0][j] * fw[0] + R[1][j] * fw[1] + R[2][j] * fw[2];
        }
        gx = gx + (c[l] * rb[1]) * fb[2] - (c[l] * rb[2]) * fb[1];
        gy = gy + (c[l] * rb[2]) * fb[0] - (c[l] * rb[0]) * fb[2];
        gz = gz + (c[l] * rb[0]) * fb[1] - (c[l] * rb[1]) * fb[0];
        for (int j = 0; j < 3; ++j) fsum[j] = fsum[j] + c[l] * fw[j];
    }
    tx = tx + gx; ty = ty + gy; tz = tz + gz;
    xn[6] = wx + (kJxx * tx + kJxy * ty + kJxz * tz) * dt;
    xn[7] = wy + (kJxy * tx + kJyy * ty + kJyz * tz) * dt;
    xn[8] = wz + (kJxz * tx + kJyz * ty + kJzz * tz) * dt;
    xn[9] = x[9] + (fsum[0] / kMass) * dt;
    xn[10] = x[10] + (fsum[1] / kMass) * dt;
    xn[11] = x[11] + (kGrav + fsum[2] / kMass) * dt;
    for (int l = 0; l < 4; ++l)
        for (int j = 0; j < 3; ++j)
            xn[12 + 3 * l + j] = x[12 + 3 * l + j] + ((1.0 - c[l]) * u[12 + 3 * l + j]) * dt;
}

// World position of foot `leg` (0=FR,1=FL,2=HR,3=HL) from pos, eul, leg joint angles.
template <class S>
inline void foot_position(const S pos[3], const S eul[3], const S q[3], int leg, S p[3]) {
    const double side = (leg & 1) ? 1.0 : -1.0;   // (-1)^(leg+1): right legs negative y
    const double fore = (leg < 2) ? 1.0 : -1.0;   // (-1)^floor((leg+1)/3)
    S q1 = q[0], q2 = -q[1], q3 = -q[2];
    S s1 = sin(q1), c1 = cos(q1), s2 = sin(q2), c2 = cos(q2), s3 = sin(q3), c3 = cos(q3);
    const double l1 = kAbad * side;
    S pb[3];
    pb[0] = kHipX * fore + (kShank * (c2 * s3 + s2 * c3) + kThigh * s2);
    pb[1] = kHipY * side + (kShank * (s1 * s2 * s3 - s1 * c2 * c3) + (c1 * l1 - s1 * (kThigh * c2)));
    pb[2] = kShank * (c1 * c2 * c3 - c1 * s2 * s3) + (c1 * (kThigh * c2) + s1 * l1);
    S R[3][3];
    rotation_zyx(eul[0], eul[1], eul[2], R);
    for (int i = 0; i < 3; ++i) p[i] = R[i][0] * pb[0] + R[i][1] * pb[1] + R[i][2] * pb[2] + pos[i];
}

// ---- dense outputs in the layout the reference's callers see ----

inline void dynamics(const double x[24], const double u[24], double dt, const int contact[4], double xn[24]) {
    double c[4] = {(double)contact[0], (double)contact[1], (double)contact[2], (double)contact[3]};
    hkd_step<double>(x, u, dt, c, xn);
}

// A, B column-major 24x24 (HKDModel.h:46-61: zeroed, then the CCS scatter).
inline void dynamics_partial(const double x[24], const double u[24], double dt, const int contact[4],
                             double A[576], double B[576]) {
    using D = Dual<48>;
    static thread_local D xd[24], ud[24], xn[24];
    for (int i = 0; i < 24; ++i) { xd[i] = D(x[i]); xd[i].d[i] = 1.0; ud[i] = D(u[i]); ud[i].d[24 + i] = 1.0; }
    double c[4] = {(double)contact[0], (double)contact[1], (double)contact[2], (double)contact[3]};
    hkd_step<D>(xd, ud, dt, c, xn);
    for (int j = 0; j < 24; ++j)
        for (int i = 0; i < 24; ++i) { A[i + 24 * j] = xn[i].d[j]; B[i + 24 * j] = xn[i].d[24 + j]; }
}

// J column-major 3x18 with columns [pos(3) eul(3) qJ(12)] (HKDReset.h:100-127).
inline void foot_jacobian(const double pos[3], const double eul[3], const double q[3], int leg, double J[54]) {
    using D = Dual<9>;
    D pd[3], ed[3], qd[3], p[3];
    for (int i = 0; i < 3; ++i) { pd[i] = D(pos[i]); pd[i].d[i] = 1; ed[i] = D(eul[i]); ed[i].d[3 + i] = 1; qd[i] = D(q[i]); qd[i].d[6 + i] = 1; }
    foot_position<D>(pd, ed, qd, leg, p);
    for (int k = 0; k < 54; ++k) J[k] = 0.0;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 6; ++j) J[i + 3 * j] = p[i].d[j];
        for (int j = 0; j < 3; ++j) J[i + 3 * (6 + 3 * leg + j)] = p[i].d[6 + j];
    }
}

}  // namespace hkd_port
