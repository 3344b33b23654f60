"""CPU tests: the EKF oracle (oracle/orc_ekf.c) against the committed cv2 transcription of the reference's
kalmanClass.cpp (tests/golden/gen_ekf_golden.py), the H matrix against a numeric Jacobian like the reference's
own getHNumeric (kalmanClass.cpp:1327-1377, JAC_EPS 1e-4), and the gate definition's invariants."""
import os

import numpy as np
import pytest

from oracle_lib import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["ekf_curves4", "ekf_curves2_points3", "ekf_points_only"]


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name", CASES)
def test_update_matches_cv2_golden(name):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    o = oracle()
    if name == "ekf_points_only":
        x, P = o.ekf_update_points(g["x"], g["P"], g["point_meas"], g["point_col"])
    else:
        x, P, H, S, K = o.ekf_update_curves_points(g["x"], g["P"], g["z"], g["A"], g["curve_col"], g["point_meas"],
                                                   g["point_col"], want_hsk=True)
        assert np.array_equal(H, g["H"])  # H needs no matrix inverse: identical
        assert rel(S, g["S"]) < 1e-14 and rel(K, g["K"]) < 1e-11
    assert rel(x, g["x_upd"]) < 1e-12 and rel(P, g["P_upd"]) < 1e-11
    assert np.array_equal(P, P.T)  # 0.5*(P + P^T) is exactly symmetric


@pytest.mark.parametrize("name", CASES)
def test_predict_matches_cv2_golden(name):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    x, P = oracle().ekf_predict(g["x"], g["P"])
    assert rel(x, g["x_pred"]) < 1e-15 and rel(P, g["P_pred"]) < 1e-14


def test_split_matrices():
    import ctypes as C
    g = np.load(os.path.join(HERE, "golden", "ekf_curves4.npz"))
    A1, A2 = np.zeros(16), np.zeros(16)
    dp = C.POINTER(C.c_double)
    oracle().lib.orc_ekf_split_matrices(float(g["split_t"]), A1.ctypes.data_as(dp), A2.ctypes.data_as(dp))
    assert np.allclose(A1.reshape(4, 4), g["A1"], rtol=1e-15) and np.allclose(A2.reshape(4, 4), g["A2"], rtol=1e-15)
    # splitting and re-joining a cubic reproduces its end points
    assert A1[0] == 1.0 and A2[15] == 1.0 and np.allclose(A1.reshape(4, 4)[3], A2.reshape(4, 4)[0])


def test_h_matches_numeric_jacobian_of_predicted_measurement():
    """Central differences of the predicted curve measurement w.r.t. the state reproduce the curve rows of H."""
    import ctypes as C
    g = np.load(os.path.join(HERE, "golden", "ekf_curves4.npz"))
    o = oracle()
    _, _, H, _, _ = o.ekf_update_curves_points(g["x"], g["P"], g["z"], g["A"], g["curve_col"], want_hsk=True)
    dp = C.POINTER(C.c_double)
    o.lib.orc_ekf_predicted_curve.argtypes = [dp, dp, C.c_int, dp]
    x0 = g["x"].copy()
    eps = 1e-4
    for k, col in enumerate(g["curve_col"]):
        A = np.ascontiguousarray(g["A"][k])
        for s in list(range(6)) + list(range(col, col + 8)):
            zp, zm = np.zeros(8), np.zeros(8)
            xp, xm = x0.copy(), x0.copy()
            xp[s] += eps
            xm[s] -= eps
            o.lib.orc_ekf_predicted_curve(xp.ctypes.data_as(dp), A.ctypes.data_as(dp), int(col), zp.ctypes.data_as(dp))
            o.lib.orc_ekf_predicted_curve(xm.ctypes.data_as(dp), A.ctypes.data_as(dp), int(col), zm.ctypes.data_as(dp))
            assert np.allclose((zp - zm) / (2 * eps), H[8 * k:8 * k + 8, s], atol=1e-6)


def test_pinv_svd_is_an_inverse():
    import ctypes as C
    rng = np.random.default_rng(0)
    dp = C.POINTER(C.c_double)
    for n in (1, 3, 35, 65):
        G = rng.standard_normal((n, n + 3))
        S = G @ G.T + np.eye(n)
        Si = np.zeros((n, n))
        assert oracle().lib.orc_pinv_svd(S.ctypes.data_as(dp), Si.ctypes.data_as(dp), n) == 0
        assert np.abs(S @ Si - np.eye(n)).max() < 1e-11


def _gate_case(seed=0, L=40, nz=25):
    rng = np.random.default_rng(seed)
    N = 12 + 3 * L
    G = rng.standard_normal((N, 16))
    P = G @ G.T / 16 * 0.05 + 0.25 * np.eye(N)
    P = 0.5 * (P + P.T)
    x = np.zeros(N)
    x[:6] = (0.2, -0.1, -1.0, 0.01, -0.02, 0.3)
    x[12:] = rng.uniform(-50, 50, N - 12)
    cols = 12 + 3 * np.arange(L, dtype=np.int32)
    return x, P, cols, rng


def test_gate_definition_invariants():
    x, P, cols, rng = _gate_case()
    o = oracle()
    pack = o.gate_pack(x, P, cols)
    # measurements exactly at the predicted position of landmark j gate to j with d2 = 0
    z = pack[:, :3].copy()
    idx, d2 = o.gate(x, P, cols, z, 11.345)
    assert np.array_equal(idx, np.arange(len(cols))) and np.all(d2 == 0.0)
    # far-away clutter gates to -1; duplicates of a landmark keep the lowest index
    idx, d2 = o.gate(x, P, cols, np.full((3, 3), 1e4), 11.345)
    assert np.all(idx == -1) and np.all(np.isinf(d2))
    x2 = x.copy()
    x2[12 + 3 * 7: 12 + 3 * 8] = x2[12:15]
    P2 = P.copy()
    a, b = slice(12, 15), slice(12 + 21, 12 + 24)
    P2[b, :] = P2[a, :]
    P2[:, b] = P2[:, a]
    P2[b, b] = P2[a, a]
    pack2 = o.gate_pack(x2, P2, cols)
    idx, _ = o.gate(x2, P2, cols, pack2[7:8, :3] + 0.01, 11.345)
    assert idx[0] == 0
    # multi-threaded sweep gives the same answer
    zz = rng.uniform(-50, 50, (200, 3))
    i1, d1 = o.gate(x, P, cols, zz, 50.0, nthreads=1)
    i4, d4 = o.gate(x, P, cols, zz, 50.0, nthreads=4)
    assert np.array_equal(i1, i4) and np.array_equal(d1, d4)


def test_augmentation_matches_cv2_golden():
    """AddFirstStates / AddNewCurve / AddNewPoints (kalmanClass.cpp:275-666) against the cv2 transcription."""
    import ctypes as C
    g = np.load(os.path.join(HERE, "golden", "ekf_augment.npz"))
    L = oracle().lib
    dp = C.POINTER(C.c_double)
    L.orc_ekf_add_new_curve.restype = C.c_int
    L.orc_ekf_add_new_points.restype = C.c_int
    ld = 64
    x, P = np.zeros(ld), np.zeros((ld, ld))
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(dp)
    L.orc_ekf_add_first_states(x.ctypes.data_as(dp), P.ctypes.data_as(dp), ld, c(g["z19"]))
    assert np.array_equal(x[:28], g["x0"]) and np.array_equal(P[:28, :28], g["P0"])
    x[:28], P[:28, :28] = g["x_in"], g["P_in"]
    N = L.orc_ekf_add_new_curve(x.ctypes.data_as(dp), P.ctypes.data_as(dp), ld, 28, c(g["z8"]), c(g["A"]))
    assert N == 36 and rel(x[:N], g["x1"]) < 1e-13 and rel(P[:N, :N], g["P1"]) < 1e-13
    N = L.orc_ekf_add_new_points(x.ctypes.data_as(dp), P.ctypes.data_as(dp), ld, N, c(g["pts"]), 3)
    assert N == 45 and rel(x[:N], g["x2"]) < 1e-13 and rel(P[:N, :N], g["P2"]) < 1e-13
    assert np.all(P[36:39, 39:45] == 0.0)  # points added together get no cross-covariance (reference behaviour)
