// ORACLE (test infrastructure only — never linked by the product).
//
// Small fixed-size dense kernels that restate the Eigen calls the reference's
// Riccati recursion relies on (HSDDPSolver/source/SinglePhase.cpp:320-358):
//   * dense products evaluated left-to-right, plain sum over the inner index
//   * Eigen::LDLT<MatrixXd>::compute(...).isPositive()   (typedef HSDDP_CPPTypes.h:63-64)
//   * fixed-size 24x24 .inverse()  ==  PartialPivLU, then solve against identity
// Eigen itself is an un-vendored, unpinned dependency of the reference and is
// NOT present in this image, so these follow Eigen's published
// algorithms (>= 3.3): LDLT = in-place lower, diagonal pivoting on max |a_ii|,
// sign tracking; PartialPivLU = row pivoting on max |a_ik|.  "parity unpinned":
// bit-level agreement with an Eigen build cannot be checked here.
#pragma once
#include <cmath>
#include <cstring>

namespace oracle {

constexpr int NX = 24;
constexpr int NU = 24;

struct Vec24 {
    double v[24];
    double& operator[](int i) { return v[i]; }
    const double& operator[](int i) const { return v[i]; }
    void zero() { std::memset(v, 0, sizeof v); }
};

// column-major 24x24
struct Mat24 {
    double m[576];
    double& operator()(int i, int j) { return m[i + 24 * j]; }
    const double& operator()(int i, int j) const { return m[i + 24 * j]; }
    void zero() { std::memset(m, 0, sizeof m); }
    void identity() { zero(); for (int i = 0; i < 24; ++i) m[i * 25] = 1.0; }
};

// C = A * B
inline void matmul(const Mat24& A, const Mat24& B, Mat24& C) {
    for (int j = 0; j < 24; ++j) {
        double col[24];
        for (int i = 0; i < 24; ++i) col[i] = 0.0;
        for (int k = 0; k < 24; ++k) {
            const double b = B(k, j);
            const double* a = &A.m[24 * k];
            for (int i = 0; i < 24; ++i) col[i] += a[i] * b;
        }
        for (int i = 0; i < 24; ++i) C(i, j) = col[i];
    }
}
// C = A^T * B
inline void matmul_tn(const Mat24& A, const Mat24& B, Mat24& C) {
    for (int j = 0; j < 24; ++j)
        for (int i = 0; i < 24; ++i) {
            double s = 0.0;
            const double* a = &A.m[24 * i];
            const double* b = &B.m[24 * j];
            for (int k = 0; k < 24; ++k) s += a[k] * b[k];
            C(i, j) = s;
        }
}
// y = A * x
inline void matvec(const Mat24& A, const Vec24& x, Vec24& y) {
    double acc[24];
    for (int i = 0; i < 24; ++i) acc[i] = 0.0;
    for (int k = 0; k < 24; ++k) {
        const double xk = x[k];
        const double* a = &A.m[24 * k];
        for (int i = 0; i < 24; ++i) acc[i] += a[i] * xk;
    }
    for (int i = 0; i < 24; ++i) y[i] = acc[i];
}
// y = A^T * x
inline void matvec_t(const Mat24& A, const Vec24& x, Vec24& y) {
    for (int i = 0; i < 24; ++i) {
        double s = 0.0;
        const double* a = &A.m[24 * i];
        for (int k = 0; k < 24; ++k) s += a[k] * x[k];
        y[i] = s;
    }
}
inline double dot(const Vec24& a, const Vec24& b) {
    double s = 0.0;
    for (int i = 0; i < 24; ++i) s += a[i] * b[i];
    return s;
}

// Eigen::LDLT sign test.  Returns true iff LDLT(M).isPositive(), i.e. the
// pivoted factorisation met no negative pivot.  Only the lower triangle of M
// is read (Eigen's default UpLo = Lower).  n <= 24.
inline bool ldlt_is_positive(const Mat24& Min, int n = 24) {
    double a[24][24];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) a[i][j] = Min(i, j);
    enum { ZeroSign, PositiveSemiDef, NegativeSemiDef, Indefinite } sign = ZeroSign;
    double temp[24];
    for (int k = 0; k < n; ++k) {
        // largest remaining diagonal entry
        int piv = k;
        double big = std::fabs(a[k][k]);
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(a[i][i]) > big) { big = std::fabs(a[i][i]); piv = i; }
        if (piv != k) {
            // symmetric swap of rows/cols k and piv, lower triangle only
            for (int j = 0; j < k; ++j) std::swap(a[k][j], a[piv][j]);
            for (int i = piv + 1; i < n; ++i) std::swap(a[i][k], a[i][piv]);
            std::swap(a[k][k], a[piv][piv]);
            for (int i = k + 1; i < piv; ++i) std::swap(a[i][k], a[piv][i]);
        }
        const int rs = n - k - 1;
        if (k > 0) {
            for (int j = 0; j < k; ++j) temp[j] = a[j][j] * a[k][j];
            double s = 0.0;
            for (int j = 0; j < k; ++j) s += a[k][j] * temp[j];
            a[k][k] -= s;
            for (int i = k + 1; i < n; ++i) {
                double t = 0.0;
                for (int j = 0; j < k; ++j) t += a[i][j] * temp[j];
                a[i][k] -= t;
            }
        }
        const double akk = a[k][k];
        const bool valid = std::fabs(akk) > 0.0;
        if (k == 0 && !valid) return true;  // zero matrix: ZeroSign counts as positive
        if (rs > 0 && valid)
            for (int i = k + 1; i < n; ++i) a[i][k] /= akk;
        if (sign == PositiveSemiDef) { if (akk < 0) sign = Indefinite; }
        else if (sign == NegativeSemiDef) { if (akk > 0) sign = Indefinite; }
        else if (sign == ZeroSign) { if (akk > 0) sign = PositiveSemiDef; else if (akk < 0) sign = NegativeSemiDef; }
    }
    return sign == PositiveSemiDef || sign == ZeroSign;
}

// Inverse by LU with partial (row) pivoting, then forward/back substitution
// against the permuted identity (Eigen PartialPivLU::inverse()).
inline void inverse_partial_piv_lu(const Mat24& Min, Mat24& Inv) {
    constexpr int n = 24;
    double lu[24][24];
    int perm[24];
    for (int i = 0; i < n; ++i) { perm[i] = i; for (int j = 0; j < n; ++j) lu[i][j] = Min(i, j); }
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double big = std::fabs(lu[k][k]);
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(lu[i][k]) > big) { big = std::fabs(lu[i][k]); piv = i; }
        if (piv != k) {
            for (int j = 0; j < n; ++j) std::swap(lu[k][j], lu[piv][j]);
            std::swap(perm[k], perm[piv]);
        }
        if (lu[k][k] != 0.0) {
            for (int i = k + 1; i < n; ++i) lu[i][k] /= lu[k][k];
        }
        for (int i = k + 1; i < n; ++i) {
            const double lik = lu[i][k];
            for (int j = k + 1; j < n; ++j) lu[i][j] -= lik * lu[k][j];
        }
    }
    // solve L U X = P I, column by column
    for (int c = 0; c < n; ++c) {
        double y[24];
        for (int i = 0; i < n; ++i) y[i] = (perm[i] == c) ? 1.0 : 0.0;
        for (int i = 0; i < n; ++i) {
            double s = y[i];
            for (int j = 0; j < i; ++j) s -= lu[i][j] * y[j];
            y[i] = s;
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < n; ++j) s -= lu[i][j] * y[j];
            y[i] = s / lu[i][i];
        }
        for (int i = 0; i < n; ++i) Inv(i, c) = y[i];
    }
}

}  // namespace oracle
