// ORACLE (test infrastructure only — never linked by the product).
//
// The model-independent sweeps of SinglePhase<T, xs, us, ys> for ANY instantiation — the reference
// instantiates <24,24,0>, <12,12,0> and <36,12,12> (HSDDPSolver/source/SinglePhase.cpp:538-540) — restated with
// run-time sizes: backward_sweep incl. the ys > 0 output terms (SinglePhase.cpp:299-367, C/D terms :329-336) and
// linear_rollout (SinglePhase.cpp:145-178).  The reference ships a model, costs and a problem only for <24,24,0>
// (the HKD path, hsddp_oracle.cpp); for the other two the inputs are what LQ_approximation would have left in the
// phase's storage: A, B, C, D, RCostData {lx, lu, ly, lxx, luu, lux, lyy}, TCostData {Phix, Phixx}, Defect.
// Same Eigen semantics as linalg.hpp (products left to right, pivoted LDLT sign test, partial-pivot LU inverse);
// at <24,24,0> the results are bit-identical to Problem::phase_backward_sweep / phase_linear_rollout
// (tests/test_generic_phase.py).  "parity unpinned" at the solver level, like the rest of the oracle.
//
// All matrices column-major (Eigen default): A xs*xs, B xs*us, C ys*xs, D ys*us, lxx xs*xs, luu us*us,
// lux us*xs, lyy ys*ys, K us*xs, H xs*xs.
#pragma once
#include <cmath>
#include <cstddef>
#include <utility>
#include <vector>

namespace oracle {
namespace generic {

typedef std::vector<double> V;

// C (m x n) = A (m x k) * B (k x n): plain sum over the inner index, in order
inline void mm(int m, int k, int n, const double* A, const double* B, double* C) {
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) C[i + (size_t)m * j] = 0.0;
    for (int j = 0; j < n; ++j)
        for (int l = 0; l < k; ++l) {
            const double b = B[l + (size_t)k * j];
            for (int i = 0; i < m; ++i) C[i + (size_t)m * j] += A[i + (size_t)m * l] * b;
        }
}
// C (m x n) = A^T * B with A (k x m), B (k x n)
inline void mm_tn(int m, int k, int n, const double* A, const double* B, double* C) {
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s += A[l + (size_t)k * i] * B[l + (size_t)k * j];
            C[i + (size_t)m * j] = s;
        }
}
inline void mv(int m, int k, const double* A, const double* x, double* y) {
    for (int i = 0; i < m; ++i) y[i] = 0.0;
    for (int l = 0; l < k; ++l)
        for (int i = 0; i < m; ++i) y[i] += A[i + (size_t)m * l] * x[l];
}
inline void mv_t(int m, int k, const double* A, const double* x, double* y) {  // y (m) = A^T x, A (k x m)
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int l = 0; l < k; ++l) s += A[l + (size_t)k * i] * x[l];
        y[i] = s;
    }
}
inline double dot(int n, const double* a, const double* b) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

// Eigen::LDLT<MatrixXd>::compute(M).isPositive(): pivoted (largest |diagonal|) in-place lower LDL^T with sign
// tracking — linalg.hpp:ldlt_is_positive with a run-time size.
inline bool ldlt_is_positive(int n, const double* M) {
    std::vector<double> a((size_t)n * n), temp(n);
    auto at = [&](int i, int j) -> double& { return a[(size_t)i * n + j]; };
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) at(i, j) = M[i + (size_t)n * j];
    enum { ZeroSign, PositiveSemiDef, NegativeSemiDef, Indefinite } sign = ZeroSign;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double big = std::fabs(at(k, k));
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(at(i, i)) > big) { big = std::fabs(at(i, i)); piv = i; }
        if (piv != k) {
            for (int j = 0; j < k; ++j) std::swap(at(k, j), at(piv, j));
            for (int i = piv + 1; i < n; ++i) std::swap(at(i, k), at(i, piv));
            std::swap(at(k, k), at(piv, piv));
            for (int i = k + 1; i < piv; ++i) std::swap(at(i, k), at(piv, i));
        }
        const int rs = n - k - 1;
        if (k > 0) {
            for (int j = 0; j < k; ++j) temp[j] = at(j, j) * at(k, j);
            double s = 0.0;
            for (int j = 0; j < k; ++j) s += at(k, j) * temp[j];
            at(k, k) -= s;
            for (int i = k + 1; i < n; ++i) {
                double t = 0.0;
                for (int j = 0; j < k; ++j) t += at(i, j) * temp[j];
                at(i, k) -= t;
            }
        }
        const double akk = at(k, k);
        const bool valid = std::fabs(akk) > 0.0;
        if (k == 0 && !valid) return true;
        if (rs > 0 && valid)
            for (int i = k + 1; i < n; ++i) at(i, k) /= akk;
        if (sign == PositiveSemiDef) { if (akk < 0) sign = Indefinite; }
        else if (sign == NegativeSemiDef) { if (akk > 0) sign = Indefinite; }
        else if (sign == ZeroSign) { if (akk > 0) sign = PositiveSemiDef; else if (akk < 0) sign = NegativeSemiDef; }
    }
    return sign == PositiveSemiDef || sign == ZeroSign;
}

// Eigen PartialPivLU::inverse() — linalg.hpp:inverse_partial_piv_lu with a run-time size
inline void inverse_partial_piv_lu(int n, const double* M, double* Inv) {
    std::vector<double> lu((size_t)n * n), y(n);
    std::vector<int> perm(n);
    auto at = [&](int i, int j) -> double& { return lu[(size_t)i * n + j]; };
    for (int i = 0; i < n; ++i) { perm[i] = i; for (int j = 0; j < n; ++j) at(i, j) = M[i + (size_t)n * j]; }
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double big = std::fabs(at(k, k));
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(at(i, k)) > big) { big = std::fabs(at(i, k)); piv = i; }
        if (piv != k) {
            for (int j = 0; j < n; ++j) std::swap(at(k, j), at(piv, j));
            std::swap(perm[k], perm[piv]);
        }
        if (at(k, k) != 0.0)
            for (int i = k + 1; i < n; ++i) at(i, k) /= at(k, k);
        for (int i = k + 1; i < n; ++i) {
            const double lik = at(i, k);
            for (int j = k + 1; j < n; ++j) at(i, j) -= lik * at(k, j);
        }
    }
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) y[i] = (perm[i] == c) ? 1.0 : 0.0;
        for (int i = 0; i < n; ++i) {
            double s = y[i];
            for (int j = 0; j < i; ++j) s -= at(i, j) * y[j];
            y[i] = s;
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < n; ++j) s -= at(i, j) * y[j];
            y[i] = s / at(i, i);
        }
        for (int i = 0; i < n; ++i) Inv[i + (size_t)n * c] = y[i];
    }
}

// Flat views of one phase's storage (Trajectory<T,xs,us,ys>, TrajectoryManagement.h:22-82).  Inputs are const
// pointers; per-stage arrays are stage-major ([N][...]), node arrays have N + 1 rows.
struct PhaseData {
    int xs, us, ys, N;
    const double *A, *B, *C, *D;                      // [N][xs*xs], [N][xs*us], [N][ys*xs], [N][ys*us]
    const double *lx, *lu, *ly, *lxx, *luu, *lux, *lyy;  // RCostData per stage
    const double *Phix, *Phixx;                       // TCostData
    const double* Defect;                             // [N+1][xs]
};

// SinglePhase::backward_sweep (SinglePhase.cpp:299-367).  Outputs: dU [N][us], K [N][us*xs], G [N+1][xs],
// H [N+1][xs*xs], dV[2].  Stages below a failed one keep whatever the caller had there (the reference breaks).
inline bool backward_sweep(const PhaseData& p, double reg, const double* Gprime, const double* Hprime,
                           double* dU, double* K, double* G, double* H, double* dV) {
    const int xs = p.xs, us = p.us, ys = p.ys, N = p.N;
    const size_t xx = (size_t)xs * xs, xu = (size_t)xs * us, uu = (size_t)us * us, yx = (size_t)ys * xs, yu = (size_t)ys * us, yy = (size_t)ys * ys;
    bool success = true;
    for (int j = 0; j < xs; ++j) G[(size_t)N * xs + j] = p.Phix[j] + Gprime[j];
    for (size_t j = 0; j < xx; ++j) H[(size_t)N * xx + j] = p.Phixx[j] + Hprime[j];
    double dV_1 = 0, dV_2 = 0;
    V AtH(xx), BtH(xu), Qxx(xx), Quu(uu), Qux(xu), Quu_s(uu), inv1(uu), Quu_inv(uu), tmp(xx), tmpu(uu), tmpux(xu), QuxT_Qi(xu);
    V CtL(yx), DtL(yu);
    V Qx(xs), Qu(us), Gnext(xs), t(xs > us ? xs : us);
    for (int k = N - 1; k >= 0; --k) {
        const double* Ak = p.A + (size_t)k * xx;
        const double* Bk = p.B + (size_t)k * xu;
        const double* Hn = H + (size_t)(k + 1) * xx;
        mv(xs, xs, Hn, p.Defect + (size_t)(k + 1) * xs, t.data());
        for (int j = 0; j < xs; ++j) Gnext[j] = G[(size_t)(k + 1) * xs + j] + t[j];
        mv_t(xs, xs, Ak, Gnext.data(), t.data()); for (int j = 0; j < xs; ++j) Qx[j] = p.lx[(size_t)k * xs + j] + t[j];
        mv_t(us, xs, Bk, Gnext.data(), t.data()); for (int j = 0; j < us; ++j) Qu[j] = p.lu[(size_t)k * us + j] + t[j];
        mm_tn(xs, xs, xs, Ak, Hn, AtH.data());
        mm_tn(us, xs, xs, Bk, Hn, BtH.data());  // us x xs
        mm(xs, xs, xs, AtH.data(), Ak, tmp.data()); for (size_t j = 0; j < xx; ++j) Qxx[j] = p.lxx[(size_t)k * xx + j] + tmp[j];
        mm(us, xs, us, BtH.data(), Bk, tmpu.data()); for (size_t j = 0; j < uu; ++j) Quu[j] = p.luu[(size_t)k * uu + j] + tmpu[j];
        mm(us, xs, xs, BtH.data(), Ak, tmpux.data()); for (size_t j = 0; j < xu; ++j) Qux[j] = p.lux[(size_t)k * xu + j] + tmpux[j];
        if (ys > 0) {  // SinglePhase.cpp:329-336
            const double* Ck = p.C + (size_t)k * yx;
            const double* Dk = p.D + (size_t)k * yu;
            const double* ly = p.ly + (size_t)k * ys;
            const double* lyy = p.lyy + (size_t)k * yy;
            mv_t(xs, ys, Ck, ly, t.data()); for (int j = 0; j < xs; ++j) Qx[j] += t[j];
            mv_t(us, ys, Dk, ly, t.data()); for (int j = 0; j < us; ++j) Qu[j] += t[j];
            mm_tn(xs, ys, ys, Ck, lyy, CtL.data());  // xs x ys
            mm_tn(us, ys, ys, Dk, lyy, DtL.data());  // us x ys
            mm(xs, ys, xs, CtL.data(), Ck, tmp.data()); for (size_t j = 0; j < xx; ++j) Qxx[j] += tmp[j];
            mm(us, ys, us, DtL.data(), Dk, tmpu.data()); for (size_t j = 0; j < uu; ++j) Quu[j] += tmpu[j];
            mm(us, ys, xs, DtL.data(), Ck, tmpux.data()); for (size_t j = 0; j < xu; ++j) Qux[j] += tmpux[j];
        }
        for (int j = 0; j < xs; ++j) Qxx[j + (size_t)xs * j] += 1.0 * reg;
        for (int j = 0; j < us; ++j) Quu[j + (size_t)us * j] += 1.0 * reg;
        Quu_s = Quu;
        for (int j = 0; j < us; ++j) Quu_s[j + (size_t)us * j] -= 1.0 * 1e-9;
        if (!ldlt_is_positive(us, Quu_s.data())) { success = false; break; }
        inverse_partial_piv_lu(us, Quu.data(), inv1.data());
        for (int j = 0; j < us; ++j)
            for (int i = 0; i < us; ++i) Quu_inv[i + (size_t)us * j] = (inv1[i + (size_t)us * j] + inv1[j + (size_t)us * i]) / 2;
        for (int j = 0; j < xs; ++j)
            for (int i = 0; i < xs; ++i) tmp[i + (size_t)xs * j] = (Qxx[i + (size_t)xs * j] + Qxx[j + (size_t)xs * i]) / 2;
        Qxx = tmp;
        mv(us, us, Quu_inv.data(), Qu.data(), t.data()); for (int j = 0; j < us; ++j) dU[(size_t)k * us + j] = -t[j];
        mm(us, us, xs, Quu_inv.data(), Qux.data(), tmpux.data()); for (size_t j = 0; j < xu; ++j) K[(size_t)k * xu + j] = -tmpux[j];
        mm_tn(xs, us, us, Qux.data(), Quu_inv.data(), QuxT_Qi.data());  // xs x us
        mv(xs, us, QuxT_Qi.data(), Qu.data(), t.data()); for (int j = 0; j < xs; ++j) G[(size_t)k * xs + j] = Qx[j] - t[j];
        mm(xs, us, xs, QuxT_Qi.data(), Qux.data(), tmp.data()); for (size_t j = 0; j < xx; ++j) H[(size_t)k * xx + j] = Qxx[j] - tmp[j];
        const double dV_k = -dot(us, Qu.data(), dU + (size_t)k * us);
        dV_1 -= dV_k;
        dV_2 += dV_k;
    }
    mv(xs, xs, H, p.Defect, t.data());  // runs even after a failed stage (SinglePhase.cpp:365)
    for (int j = 0; j < xs; ++j) G[j] += t[j];
    dV[0] = dV_1; dV[1] = dV_2;
    return success;
}

// SinglePhase::linear_rollout (SinglePhase.cpp:145-178).  dX [N+1][xs] out, dV[2] out (overwritten).
inline void linear_rollout(const PhaseData& p, double eps, const double* dx_init, const double* dU, const double* K,
                           double* dX, double* dV) {
    const int xs = p.xs, us = p.us, N = p.N;
    const size_t xx = (size_t)xs * xs, xu = (size_t)xs * us, uu = (size_t)us * us;
    double dV_1 = 0, dV_2 = 0;
    V duk(us), Kdx(us), Adx(xs), Bdu(xs), t1(xs > us ? xs : us);
    for (int j = 0; j < xs; ++j) dX[j] = dx_init[j] + eps * p.Defect[j];
    for (int k = 0; k < N; ++k) {
        const double* dxk = dX + (size_t)k * xs;
        mv(us, xs, K + (size_t)k * xu, dxk, Kdx.data());
        for (int j = 0; j < us; ++j) duk[j] = eps * dU[(size_t)k * us + j] + Kdx[j];
        mv(xs, xs, p.A + (size_t)k * xx, dxk, Adx.data());
        mv(xs, us, p.B + (size_t)k * xu, duk.data(), Bdu.data());
        for (int j = 0; j < xs; ++j) dX[(size_t)(k + 1) * xs + j] = (Adx[j] + Bdu[j]) + eps * p.Defect[(size_t)(k + 1) * xs + j];
        dV_1 += dot(xs, p.lx + (size_t)k * xs, dxk) + dot(us, p.lu + (size_t)k * us, duk.data());
        mv_t(xs, xs, p.lxx + (size_t)k * xx, dxk, t1.data());
        dV_2 += dot(xs, t1.data(), dxk);
        mv_t(us, us, p.luu + (size_t)k * uu, duk.data(), t1.data());
        dV_2 += dot(us, t1.data(), duk.data());
        mv_t(xs, us, p.lux + (size_t)k * xu, duk.data(), t1.data());  // (du^T P) . dx, counted once (Q13)
        dV_2 += dot(xs, t1.data(), dxk);
    }
    const double* dxN = dX + (size_t)N * xs;
    dV_1 += dot(xs, p.Phix, dxN);
    mv_t(xs, xs, p.Phixx, dxN, t1.data());
    dV_2 += dot(xs, t1.data(), dxN);
    dV[0] = dV_1; dV[1] = dV_2;
}

}  // namespace generic
}  // namespace oracle
