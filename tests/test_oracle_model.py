"""Oracle pinning at the model level (CPU).

The only arithmetic of the hot path that the reference ships in buildable form is
its CasADi-generated model (HKDMPC/HKD-TrajOpt/CasadiGen/source/*.cpp).  Golden
vectors generated from that code (tools/make_model_vectors.py) pin
  * the oracle's independent model port (oracle/hkd_model_port.hpp), and
  * the product's hand-derived analytic model (hkd-mpc_b200/csrc/hkd_model.cuh, host build).
"""
import os
import numpy as np
import pytest
from conftest import GOLDEN, GAIT_PATH, rel_err

TOL = 1e-13  # FP64; differences are association order of a few dozen operations


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(GOLDEN, "model_vectors.npz"))


def test_port_matches_reference_vectors(orc, vec):
    dt = float(vec["dt"])
    for i in range(vec["X"].shape[0]):
        x, u, c = vec["X"][i], vec["U"][i], vec["contact"][i]
        assert rel_err(orc.model_dynamics(orc.MODEL_PORT, x, u, dt, c), vec["Xn"][i]) < TOL
        A, B = orc.model_dynamics_partial(orc.MODEL_PORT, x, u, dt, c)
        assert rel_err(A, vec["A"][i]) < TOL and rel_err(B, vec["B"][i]) < TOL
        for leg in range(4):
            q = x[12 + 3 * leg:15 + 3 * leg]
            assert np.abs(orc.model_foot_position(orc.MODEL_PORT, x[3:6], x[0:3], q, leg) - vec["foot_pos"][i, leg]).max() < TOL
            assert np.abs(orc.model_foot_jacobian(orc.MODEL_PORT, x[3:6], x[0:3], q, leg) - vec["foot_jac"][i, leg]).max() < TOL


def test_compiled_reference_matches_its_vectors(orc, vec):
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    dt = float(vec["dt"])
    for i in range(vec["X"].shape[0]):
        x, u, c = vec["X"][i], vec["U"][i], vec["contact"][i]
        assert np.array_equal(orc.model_dynamics(orc.MODEL_REF, x, u, dt, c), vec["Xn"][i])
        A, B = orc.model_dynamics_partial(orc.MODEL_REF, x, u, dt, c)
        assert np.array_equal(A, vec["A"][i]) and np.array_equal(B, vec["B"][i])


def test_product_host_model_matches_reference_vectors(pkg, vec):
    dt = float(vec["dt"])
    for i in range(vec["X"].shape[0]):
        x, u, c = vec["X"][i], vec["U"][i], vec["contact"][i]
        assert rel_err(pkg.model_dynamics(x, u, dt, c), vec["Xn"][i]) < TOL
        A, B = pkg.model_dynamics_partial(x, u, dt, c)
        assert rel_err(A, vec["A"][i]) < TOL and rel_err(B, vec["B"][i]) < TOL
        for leg in range(4):
            q = x[12 + 3 * leg:15 + 3 * leg]
            assert np.abs(pkg.model_foot_position(x[3:6], x[0:3], q, leg) - vec["foot_pos"][i, leg]).max() < TOL
            assert np.abs(pkg.model_foot_jacobian(x[3:6], x[0:3], q, leg) - vec["foot_jac"][i, leg]).max() < TOL


def test_B_sparsity_is_the_references_60_nnz(vec):
    # hkinodyn_par_casadi.cpp:178: B has 60 structural non-zeros
    pattern = np.zeros((24, 24), bool)
    for j in range(12):
        pattern[6:9, j] = True
        pattern[9 + j % 3, j] = True
    for j in range(12, 24):
        pattern[j, j] = True
    assert pattern.sum() == 60
    assert not np.any(vec["B"][:, ~pattern])


def test_jacobians_against_finite_differences(orc):
    rng = np.random.default_rng(3)
    dt = float(np.float32(0.01))
    x = rng.normal(size=24) * 0.2
    x[5] += 0.25
    u = rng.normal(size=24) * 5
    c = np.array([1, 0, 1, 1], np.int32)
    A, B = orc.model_dynamics_partial(orc.MODEL_PORT, x, u, dt, c)
    h = 1e-6
    for j in range(24):
        e = np.zeros(24); e[j] = h
        fa = (orc.model_dynamics(orc.MODEL_PORT, x + e, u, dt, c) - orc.model_dynamics(orc.MODEL_PORT, x - e, u, dt, c)) / (2 * h)
        fb = (orc.model_dynamics(orc.MODEL_PORT, x, u + e, dt, c) - orc.model_dynamics(orc.MODEL_PORT, x, u - e, dt, c)) / (2 * h)
        assert np.abs(fa - A[:, j]).max() < 1e-7 and np.abs(fb - B[:, j]).max() < 1e-7
    for leg in range(4):
        q = x[12 + 3 * leg:15 + 3 * leg]
        J = orc.model_foot_jacobian(orc.MODEL_PORT, x[3:6], x[0:3], q, leg)
        for j in range(3):
            e = np.zeros(3); e[j] = h
            f = lambda pos, eul, qq: orc.model_foot_position(orc.MODEL_PORT, pos, eul, qq, leg)
            assert np.abs((f(x[3:6] + e, x[0:3], q) - f(x[3:6] - e, x[0:3], q)) / (2 * h) - J[:, j]).max() < 1e-7
            assert np.abs((f(x[3:6], x[0:3] + e, q) - f(x[3:6], x[0:3] - e, q)) / (2 * h) - J[:, 3 + j]).max() < 1e-7
            assert np.abs((f(x[3:6], x[0:3], q + e) - f(x[3:6], x[0:3], q - e)) / (2 * h) - J[:, 6 + 3 * leg + j]).max() < 1e-7
