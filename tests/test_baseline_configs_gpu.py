"""GPU parity at the sizes BASELINE.json quotes (configs[2]: EKF update at 1k / 5k / 20k landmarks; configs[3]: gate
10 000 measurements x 20 000 landmarks), through the C ABI (include/cps_ekf.h).

  * 5k landmarks (N = 15 012), M = 35 and M = 65: the CUDA update against the dense CPU restatement (orc_ekf.c, pinned
    to the reference's kalmanClass.cpp by tests/test_ref_pinning.py), every entry of x and P;
  * 20k landmarks (N = 60 012, P = 28.8 GB: no dense host copy exists): P = G G^T / 64 + 0.25 I is known in closed
    form, so K, S and any tile of the updated covariance P - K S K^T follow on the host in float64 from G and the
    sparse H without the matrix; >= 200 sampled 64 x 64 tiles incl. the diagonal, the pose rows and the last ragged
    tile are compared;
  * 8-state curve landmarks at N = 8 012 against the dense restatement;
  * the gate at 10 000 x 19 989 landmarks against orc_gate.c (all indices, all d2, bit for bit).
EKF criterion (kalmanClass.cpp:863-926 is the arithmetic): element-wise |a - b| <= 1e-9 * max(|b|, 1e-2 * max|P|);
both the element-wise and the max-norm relative errors are printed (run with -s) and recorded in the test's
user_properties.
"""
import ctypes as C

import numpy as np
import pytest

from curvepointslam_b200 import ekf
from oracle_lib import oracle
from parity import ekf_close

gpu = pytest.mark.gpu
dp = C.POINTER(C.c_double)
PI_REF = 3.1415926535  # common.h:114


def _measurement(rng, x, ci, pi, which_curves, A, which_points):
    """z = predicted measurement + noise (so the innovation is small, as in tracking), point measurements likewise."""
    L = oracle().lib
    L.orc_ekf_predicted_curve.argtypes = [dp, dp, C.c_int, dp]
    L.orc_ekf_rbe.argtypes = [C.c_double, C.c_double, C.c_double, dp]
    xs = np.ascontiguousarray(x)
    zh = []
    for k, c in enumerate(which_curves):
        o = np.zeros(8)
        Ak = np.ascontiguousarray(A[k])
        L.orc_ekf_predicted_curve(xs.ctypes.data_as(dp), Ak.ctypes.data_as(dp), int(ci[c]), o.ctypes.data_as(dp))
        zh.append(o)
    zhat = np.concatenate(zh + [x[2:5]])
    z = zhat + np.concatenate([rng.normal(0, 0.5, 8 * len(which_curves)), rng.normal(0, 0.05, 3)])
    R = np.zeros(9)
    L.orc_ekf_rbe(x[3], x[4], x[5], R.ctypes.data_as(dp))
    R = R.reshape(3, 3)
    pmh = np.concatenate([R @ (x[pi[j]:pi[j] + 3] - x[:3]) for j in which_points]) if len(which_points) else np.zeros(0)
    pm = pmh + rng.normal(0, 2.0, pmh.size)
    return z, zhat, pm, pmh


@gpu
@pytest.mark.parametrize("n_pts", [0, 10])
def test_update_at_5k_landmarks_vs_dense_oracle(n_pts, record_property):
    """BASELINE configs[2], 5k point landmarks: N = 15 012, n = 4 curves (M = 35) and + 10 points (M = 65)."""
    x, G, ci, pi = ekf.make_synthetic_ekf(5000, d=3, seed=6)
    N = x.size
    assert N == 15012
    P = G @ G.T / 64.0 + 0.25 * np.eye(N)
    P = 0.5 * (P + P.T)
    kf = ekf.KalmanFilter(capacity=N)
    kf.set_state(x, P, ci, pi)
    rng = np.random.default_rng(2)
    t1, t2 = kf.GetSplitMatrices(0.4)
    A = np.stack([t1, t2, np.eye(4), np.eye(4)])
    pts = np.sort(rng.choice(len(pi), n_pts, replace=False)) if n_pts else np.zeros(0, dtype=np.int64)
    z, _, pm, _ = _measurement(rng, x, ci, pi, [0, 1, 2, 3], A, pts)
    kf.PredictKF()
    kf.UpdateNCurvesAndPoints(z, 4, A, [0, 1, 2, 3], pm, pts, n_pts)
    xg, Pg = kf.getState(), kf.getCov()
    kf.close()
    xo, Po = oracle().ekf_predict(x, P)
    xo, Po = oracle().ekf_update_curves_points(xo, Po, z, A, ci, pm, pi[pts] if n_pts else None)
    ex, mx = ekf_close(xg, xo, "x")
    eP, mP = ekf_close(Pg, Po, "P")
    assert np.array_equal(Pg, Pg.T)
    record_property("x_elementwise_rel", ex); record_property("x_maxnorm_rel", mx)
    record_property("P_elementwise_rel", eP); record_property("P_maxnorm_rel", mP)
    print(f"\n5k landmarks M={35 + 3 * n_pts}: x elementwise {ex:.2e} maxnorm {mx:.2e}; P elementwise {eP:.2e} maxnorm {mP:.2e}")


def _host_update_lowrank(x, G, scale, diag, ci, pi, which_curves, A, z, which_points, pm):
    """Independent float64 evaluation of kalmanClass.cpp:863-926 for P = scale G G^T + diag I WITHOUT forming P:
    H, S and the innovation come from the dense restatement run on the few states H touches (pose + measured
    landmarks: H is zero elsewhere); T = P H^T, K = T S^-1 and any block of P - K S K^T follow from G."""
    N = x.size
    cols = list(range(12))
    for c in which_curves:
        cols += list(range(ci[c], ci[c] + 8))
    for j in which_points:
        cols += list(range(pi[j], pi[j] + 3))
    cols = np.array(cols)
    xr = x[cols]
    Pr = scale * (G[cols] @ G[cols].T) + diag * np.eye(cols.size)
    Pr = 0.5 * (Pr + Pr.T)
    cc = 12 + 8 * np.arange(len(which_curves), dtype=np.int32)
    pc = 12 + 8 * len(which_curves) + 3 * np.arange(len(which_points), dtype=np.int32)
    xr_new, _, H, S, _ = oracle().ekf_update_curves_points(xr, Pr, z, A, cc, pm, pc if len(pc) else None, want_hsk=True)
    M = H.shape[0]
    Ht = np.zeros((N, M))
    Ht[cols] = H.T
    T = scale * (G @ (G[cols].T @ H.T)) + diag * Ht
    Sinv = np.linalg.inv(S)
    K = T @ Sinv
    # innovation: the reduced update moved the reduced state by K_red nu; recover nu from the three direct rows and
    # the general relation S^-1-free:  K_red = T[cols] S^-1  =>  solve in the least-squares sense (cols.size >= M here)
    Kr = K[cols]
    dxr = xr_new - xr
    dxr[3:6] = np.arctan2(np.sin(dxr[3:6]), np.cos(dxr[3:6]))  # undo a possible wrap (none for small innovations)
    nu = np.linalg.lstsq(Kr, dxr, rcond=None)[0]
    x_new = x + K @ nu
    for a in (3, 4, 5):
        while x_new[a] > PI_REF:
            x_new[a] -= 2 * PI_REF
        while x_new[a] < -PI_REF:
            x_new[a] += 2 * PI_REF
    KS = K @ S

    def block(r0, c0, nr, nc):
        blk = scale * (G[r0:r0 + nr] @ G[c0:c0 + nc].T) - KS[r0:r0 + nr] @ K[c0:c0 + nc].T
        rr, cc2 = np.arange(r0, r0 + nr)[:, None], np.arange(c0, c0 + nc)[None, :]
        return blk + diag * (rr == cc2)

    return x_new, block, K


@gpu
def test_update_at_20k_landmarks_sampled_tiles(record_property):
    """BASELINE configs[2], 20k point landmarks: N = 60 012, P = 28.8 GB resident on the GPU, M = 35."""
    L = 20000
    x, G, ci, pi = ekf.make_synthetic_ekf(L, d=3, seed=9)
    N = x.size
    assert N == 60012
    kf = ekf.KalmanFilter(capacity=N)
    kf.set_state(x, None, ci, pi)
    kf.set_cov_lowrank(G, 1.0 / 64.0, 0.25)
    rng = np.random.default_rng(3)
    t1, t2 = kf.GetSplitMatrices(0.6)
    A = np.stack([np.eye(4), t1, t2, np.eye(4)])
    z, _, pm, _ = _measurement(rng, x, ci, pi, [0, 1, 2, 3], A, [])
    kf.UpdateNCurvesAndPoints(z, 4, A, [0, 1, 2, 3])
    xg = kf.getState()
    x_host, block, K = _host_update_lowrank(x, G, 1.0 / 64.0, 0.25, ci, pi, [0, 1, 2, 3], A, z, [], pm)
    ex, mx = ekf_close(xg, x_host, "x")
    nt = (N + 63) // 64
    last = nt - 1
    tiles = {(0, 0), (0, 1), (1, 0), (0, last), (last, 0), (last, last), (last - 1, last), (nt // 2, nt // 2)}
    tiles |= {(int(i), int(i)) for i in rng.integers(0, nt, 40)}           # diagonal tiles
    tiles |= {(0, int(j)) for j in rng.integers(0, nt, 20)}                # the pose rows (all of H's columns)
    while len(tiles) < 220:
        tiles.add((int(rng.integers(0, nt)), int(rng.integers(0, nt))))
    pmax = max(float(np.abs(block(64 * i, 64 * i, 64, 64)).max()) for i in (0, 1, nt // 2))
    worst_e = worst_m = 0.0
    for (I, J) in sorted(tiles):
        r0, c0 = 64 * I, 64 * J
        nr, nc = min(64, N - r0), min(64, N - c0)
        g = kf.getCovBlock(r0, c0, nr, nc)
        h = block(r0, c0, nr, nc)
        e, m = ekf_close(g, h, f"P tile ({I},{J})", scale=pmax)
        worst_e, worst_m = max(worst_e, e), max(worst_m, float(np.abs(g - h).max() / pmax))
        gt = kf.getCovBlock(c0, r0, nc, nr)
        assert np.array_equal(g, gt.T), f"tile ({I},{J}) is not the exact mirror of ({J},{I})"
    assert (N - 64 * last) == 44  # the last tile is ragged
    kf.close()
    record_property("tiles", len(tiles)); record_property("P_elementwise_rel", worst_e); record_property("P_maxnorm_rel", worst_m)
    record_property("x_elementwise_rel", ex)
    print(f"\n20k landmarks: {len(tiles)} tiles, P elementwise {worst_e:.2e} maxnorm {worst_m:.2e}; x elementwise {ex:.2e} maxnorm {mx:.2e}")


def test_host_lowrank_evaluation_agrees_with_dense_oracle_small():
    """The closed-form host evaluation used at 20k landmarks is itself checked against the dense restatement where
    both exist (N = 1 512)."""
    x, G, ci, pi = ekf.make_synthetic_ekf(500, d=3, seed=4)
    N = x.size
    P = G @ G.T / 64.0 + 0.25 * np.eye(N)
    P = 0.5 * (P + P.T)
    rng = np.random.default_rng(1)
    A = np.stack([np.eye(4)] * 4)
    pts = np.array([3, 77, 240])
    z, _, pm, _ = _measurement(rng, x, ci, pi, [0, 1, 2, 3], A, pts)
    xo, Po = oracle().ekf_update_curves_points(x, P, z, A, ci, pm, pi[pts])
    xh, block, _ = _host_update_lowrank(x, G, 1.0 / 64.0, 0.25, ci, pi, [0, 1, 2, 3], A, z, pts, pm)
    ekf_close(xh, xo, "x host closed form")
    ekf_close(block(0, 0, N, N), Po, "P host closed form")


@gpu
def test_update_8_state_curve_landmarks_at_1k(record_property):
    """BASELINE configs[2] with 8-state curve landmarks: N = 8 012, four of them measured through split matrices."""
    x, G, ci, pi = ekf.make_synthetic_ekf(1000, d=8, seed=8)
    N = x.size
    assert N == 8012
    P = G @ G.T / 64.0 + 0.25 * np.eye(N)
    P = 0.5 * (P + P.T)
    kf = ekf.KalmanFilter(capacity=N)
    kf.set_state(x, P, ci, pi)
    A1, A2 = kf.GetSplitMatrices(0.35)
    A = np.stack([A1, A2, A2, A1])
    which = [3, 170, 420, 999]
    rng = np.random.default_rng(3)
    z, _, pm, _ = _measurement(rng, x, ci, pi, which, A, [])
    kf.UpdateNCurvesAndPoints(z, 4, A, which)
    xg, Pg = kf.getState(), kf.getCov()
    kf.close()
    xo, Po = oracle().ekf_update_curves_points(x, P, z, A, ci[which])
    ex, mx = ekf_close(xg, xo, "x")
    eP, mP = ekf_close(Pg, Po, "P")
    assert np.array_equal(Pg, Pg.T)
    record_property("P_elementwise_rel", eP); record_property("P_maxnorm_rel", mP)
    print(f"\n8-state x 1k: x elementwise {ex:.2e} maxnorm {mx:.2e}; P elementwise {eP:.2e} maxnorm {mP:.2e}")


@gpu
def test_gate_10k_measurements_x_20k_landmarks_bit_exact(record_property):
    """BASELINE configs[3]: 10 000 measurements against the 19 989 point landmarks of the 20k map; every index and
    every d2 bit-identical to orc_gate.c.  The covariance lives on the GPU only (28.8 GB); the oracle receives the
    pieces of it the gate reads (pose rows / columns, the 3 x 3 diagonal blocks), read back from the device."""
    L = 20000
    x, G, ci, pi = ekf.make_synthetic_ekf(L, d=3, seed=7)
    N = x.size
    kf = ekf.KalmanFilter(capacity=N)
    kf.set_state(x, None, ci, pi)
    kf.set_cov_lowrank(G, 0.05 / 64.0, 0.25)
    pose_rows = kf.getCovBlock(0, 0, 6, N)
    pose_cols = kf.getCovBlock(0, 0, N, 6)
    diag3 = np.zeros((len(pi), 3, 3))
    step = 1024
    for j0 in range(0, len(pi), step):
        j1 = min(len(pi), j0 + step)
        c0, c1 = int(pi[j0]), int(pi[j1 - 1]) + 3
        blk = kf.getCovBlock(c0, c0, c1 - c0, c1 - c0)
        for j in range(j0, j1):
            o = int(pi[j]) - c0
            diag3[j] = blk[o:o + 3, o:o + 3]
    o = oracle()
    rng = np.random.default_rng(7)
    nz = 10000
    _, _, pack = o.gate_pieces(x, pose_rows, pose_cols, diag3, pi, np.zeros((1, 3)), 11.345, want_pack=True)
    pick = rng.integers(0, len(pi), nz)
    z = pack[pick, :3] + rng.normal(0, 5.0, (nz, 3))
    clutter = rng.random(nz) < 0.2
    z[clutter] = rng.uniform(-50, 50, (int(clutter.sum()), 3)) + rng.choice([-300.0, 300.0], (int(clutter.sum()), 3))
    idx, d2 = kf.gate(z, 11.345)
    io, do = o.gate_pieces(x, pose_rows, pose_cols, diag3, pi, z, 11.345, nthreads=16)
    kf.close()
    assert len(pi) == 19989
    assert np.array_equal(idx, io), f"{int((idx != io).sum())} of {nz} indices differ"
    assert np.array_equal(d2.view(np.int64), do.view(np.int64))
    assert 0.3 * nz < (idx >= 0).sum() < nz and (idx < 0).sum() > 0.1 * nz
    record_property("pairs", nz * len(pi)); record_property("associated", int((idx >= 0).sum()))
