// ORACLE (test infrastructure only — never linked by the product).
//
// Model callbacks of the HKD problem as the reference's solver sees them:
//   HKD::Model::dynamics / dynamics_partial      HKDMPC/HKD-TrajOpt/HKDModel.h:33-61
//   compute_hkd_state                            HKDMPC/HKD-TrajOpt/HKDModel.h:65-96
//   foot position / Jacobian as used by          HKDReset.h:41-136, HKDConstraints.cpp:69-171
// with the dense-scatter semantics of common/casadi_interface.cpp:46-68
// (outputs are zero-filled by the caller, then the CCS non-zeros are written).
//
// Two interchangeable back ends:
//   kRef  : the reference's own CasADi C, compiled UNMODIFIED into
//           oracle/_ref/libhkd_casadi_ref.so by oracle/Makefile and dlopen'ed here
//   kPort : hkd_model_port.hpp (independent restatement, dual-number Jacobians)
#pragma once
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <string>
#include "hkd_model_port.hpp"

namespace oracle {

typedef long long int casadi_int_t;
typedef int (*casadi_fn)(const double**, double**, casadi_int_t*, double*, int);
typedef const casadi_int_t* (*casadi_sp_fn)(casadi_int_t);

enum ModelKind { kRef = 0, kPort = 1 };

struct CasadiRef {
    void* handle = nullptr;
    casadi_fn hkinodyn = nullptr, hkinodyn_par = nullptr, foot_pos = nullptr, foot_jac[4] = {nullptr, nullptr, nullptr, nullptr};
    // CCS pattern of output 1 (B) of hkinodyn_par, read from the library itself
    int b_rows[64];
    int b_cols[64];
    int b_nnz = 0;

    bool load(const char* path) {
        if (handle) return true;
        handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        if (!handle) return false;
        hkinodyn = (casadi_fn)dlsym(handle, "hkinodyn");
        hkinodyn_par = (casadi_fn)dlsym(handle, "hkinodyn_par");
        foot_pos = (casadi_fn)dlsym(handle, "compute_foot_position");
        const char* jn[4] = {"comp_foot_jacob_1", "comp_foot_jacob_2", "comp_foot_jacob_3", "comp_foot_jacob_4"};
        for (int i = 0; i < 4; ++i) foot_jac[i] = (casadi_fn)dlsym(handle, jn[i]);
        casadi_sp_fn sp = (casadi_sp_fn)dlsym(handle, "hkinodyn_par_sparsity_out");
        if (!hkinodyn || !hkinodyn_par || !foot_pos || !foot_jac[0] || !foot_jac[1] || !foot_jac[2] || !foot_jac[3] || !sp) {
            dlclose(handle); handle = nullptr; return false;
        }
        // decode the compressed-column pattern exactly like casadi_interface.cpp:46-68
        const casadi_int_t* s = sp(1);
        int nrow = (int)s[0], ncol = (int)s[1];
        const casadi_int_t* colind = s + 2;
        const casadi_int_t* row = colind + ncol + 1;
        b_nnz = (int)colind[ncol];
        if (nrow != 24 || ncol != 24 || b_nnz > 64) { dlclose(handle); handle = nullptr; return false; }
        int nz = 0;
        for (int c = 0; c < ncol; ++c)
            while (nz < colind[c + 1]) { b_rows[nz] = (int)row[nz]; b_cols[nz] = c; ++nz; }
        // output 0 (A) must be dense 24x24
        const casadi_int_t* s0 = sp(0);
        if (s0[0] != 24 || s0[1] != 24 || s0[2 + 24] != 576) { dlclose(handle); handle = nullptr; return false; }
        return true;
    }
};

inline CasadiRef& casadi_ref() { static CasadiRef r; return r; }

struct Model {
    ModelKind kind = kPort;

    void dynamics(const double x[24], const double u[24], double dt, const int contact[4], double xn[24]) const {
        if (kind == kRef) {
            double c[4] = {(double)contact[0], (double)contact[1], (double)contact[2], (double)contact[3]};
            const double* arg[4] = {x, u, &dt, c};
            double* res[1] = {xn};
            casadi_ref().hkinodyn(arg, res, nullptr, nullptr, 0);
        } else {
            hkd_port::dynamics(x, u, dt, contact, xn);
        }
    }

    void dynamics_partial(const double x[24], const double u[24], double dt, const int contact[4],
                          double A[576], double B[576]) const {
        if (kind == kRef) {
            double c[4] = {(double)contact[0], (double)contact[1], (double)contact[2], (double)contact[3]};
            const double* arg[4] = {x, u, &dt, c};
            double bnz[64];
            double* res[2] = {A, bnz};  // A is dense column-major: written in place
            casadi_ref().hkinodyn_par(arg, res, nullptr, nullptr, 0);
            std::memset(B, 0, 576 * sizeof(double));
            const CasadiRef& r = casadi_ref();
            for (int k = 0; k < r.b_nnz; ++k) B[r.b_rows[k] + 24 * r.b_cols[k]] = bnz[k];
        } else {
            hkd_port::dynamics_partial(x, u, dt, contact, A, B);
        }
    }

    void foot_position(const double pos[3], const double eul[3], const double q[3], int leg, double p[3]) const {
        if (kind == kRef) {
            double id = (double)leg + 1.0;
            const double* arg[4] = {pos, eul, q, &id};
            double* res[1] = {p};
            casadi_ref().foot_pos(arg, res, nullptr, nullptr, 0);
        } else {
            hkd_port::foot_position<double>(pos, eul, q, leg, p);
        }
    }

    // J: 3x18 column-major, columns [pos eul qJ(12)]
    void foot_jacobian(const double pos[3], const double eul[3], const double q[3], int leg, double J[54]) const {
        if (kind == kRef) {
            const double* arg[3] = {pos, eul, q};
            double* res[1] = {J};
            casadi_ref().foot_jac[leg](arg, res, nullptr, nullptr, 0);
        } else {
            hkd_port::foot_jacobian(pos, eul, q, leg, J);
        }
    }

    // compute_hkd_state, HKDModel.h:65-96: joint angles for swing legs, 3-D foot position for stance legs.
    void hkd_state(const double eul[3], const double pos[3], const double qJ[12], const int contact[4], double qdummy[12]) const {
        for (int l = 0; l < 4; ++l) {
            if (contact[l] == 0) {
                for (int j = 0; j < 3; ++j) qdummy[3 * l + j] = qJ[3 * l + j];
            } else {
                foot_position(pos, eul, qJ + 3 * l, l, qdummy + 3 * l);
            }
        }
    }
};

}  // namespace oracle
