"""GPU parity tests of the EKF predict/update and the Mahalanobis gate through the C ABI (include/cps_ekf.h).
EKF bar (north star): state and covariance within 1e-9 relative of the oracle / golden vectors, ELEMENT-WISE:
|a - b| <= 1e-9 * max(|b|, 1e-2 * max|b|) (tests/parity.py; entries that cancel below a hundredth of the largest one
are held to 1e-11 * max|b| absolute).  Gate bar: bit-exact indices (and bit-exact d2) against the CPU definition.
Fixtures: tests/golden/ref_cps.npz are outputs of the reference's own kalmanClass.cpp (tests/golden/gen_ref_golden.py);
tests/golden/ekf_*.npz of a cv2 transcription."""
import os

import numpy as np
import pytest

from curvepointslam_b200 import ekf
from oracle_lib import oracle
from parity import elementwise_rel

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-9


def rel(a, b):
    """element-wise relative error (tests/parity.py), not the max-norm one"""
    return elementwise_rel(a, b)


def load(name):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    n_c = {"ekf_curves4": 6, "ekf_curves2_points3": 3, "ekf_points_only": 2}[name]
    N = g["P"].shape[0]
    curve_inds = 12 + 8 * np.arange(n_c)
    point_inds = np.arange(12 + 8 * n_c, N, 3)
    return g, curve_inds, point_inds


@pytest.mark.parametrize("name", ["ekf_curves4", "ekf_curves2_points3"])
def test_update_curves_points_vs_golden_and_oracle(name):
    g, ci, pi = load(name)
    kf = ekf.KalmanFilter(capacity=g["P"].shape[0])
    kf.set_state(g["x"], g["P"], ci, pi)
    curve_num = [int(np.where(ci == c)[0][0]) for c in g["curve_col"]]
    point_num = [int(np.where(pi == c)[0][0]) for c in g["point_col"]]
    kf.UpdateNCurvesAndPoints(g["z"], len(curve_num), g["A"], curve_num, g["point_meas"], point_num, len(point_num))
    x, P = kf.getState(), kf.getCov()
    assert rel(x, g["x_upd"]) < TOL and rel(P, g["P_upd"]) < TOL      # cv2 transcription of the reference
    xo, Po = oracle().ekf_update_curves_points(g["x"], g["P"], g["z"], g["A"], g["curve_col"], g["point_meas"], g["point_col"])
    assert rel(x, xo) < TOL and rel(P, Po) < TOL                        # C restatement
    assert np.array_equal(P, P.T)                                       # exactly symmetric, like the reference
    t = kf.last_timing()
    assert t["launches"] == 3 and t["cov_kernel_ms"] > 0  # front (build + S + S^-1), rows (P H^T + gain), covariance
    kf.close()


def test_update_points_vs_golden():
    g, ci, pi = load("ekf_points_only")
    kf = ekf.KalmanFilter(capacity=64)  # capacity larger than N: padding stays inert
    kf.set_state(g["x"], g["P"], ci, pi)
    point_num = [int(np.where(pi == c)[0][0]) for c in g["point_col"]]
    kf.UpdatePoints(g["point_meas"], point_num, len(point_num))
    assert rel(kf.getState(), g["x_upd"]) < TOL and rel(kf.getCov(), g["P_upd"]) < TOL
    kf.close()


@pytest.mark.parametrize("name", ["ekf_curves4", "ekf_curves2_points3", "ekf_points_only"])
def test_predict_vs_golden(name):
    g, ci, pi = load(name)
    kf = ekf.KalmanFilter(capacity=g["P"].shape[0])
    kf.set_state(g["x"], g["P"], ci, pi)
    kf.PredictKF()
    assert rel(kf.getState(), g["x_pred"]) < TOL and rel(kf.getCov(), g["P_pred"]) < TOL
    kf.close()


def test_split_matrices_and_bad_arguments():
    g, ci, pi = load("ekf_curves4")
    kf = ekf.KalmanFilter(capacity=60)
    A1, A2 = kf.GetSplitMatrices(float(g["split_t"]))
    assert np.allclose(A1, g["A1"], rtol=1e-15) and np.allclose(A2, g["A2"], rtol=1e-15)
    kf.set_state(g["x"], g["P"], ci, pi)
    with pytest.raises(Exception):
        kf.UpdateNCurvesAndPoints(g["z"], 1, g["A"][:1], [99])          # curve index out of range
    with pytest.raises(Exception):
        kf.set_state(np.zeros(100), None)                                # larger than capacity
    kf.close()


REF = np.load(os.path.join(HERE, "golden", "ref_cps.npz"))


@pytest.mark.parametrize("name", ["c4", "c2p3", "p4", "c4p10"])
def test_predict_and_update_vs_reference_kalmanclass_fixtures(name):
    """The CUDA path against outputs of the reference's own KalmanFilter (kalmanClass.cpp compiled unchanged)."""
    pre = f"ekf_{name}_"
    x, P, ci, pi = REF[pre + "x"], REF[pre + "P"], REF[pre + "ci"], REF[pre + "pi"]
    kf = ekf.KalmanFilter(capacity=x.size)
    kf.set_state(x, P, ci, pi)
    kf.PredictKF()
    assert rel(kf.getState(), REF[pre + "x_ref_pred"]) < TOL and rel(kf.getCov(), REF[pre + "P_ref_pred"]) < TOL
    kf.set_state(x, P, ci, pi)
    cn, pn = REF[pre + "cn"], REF[pre + "pn"]
    if len(cn):
        kf.UpdateNCurvesAndPoints(REF[pre + "z"], len(cn), REF[pre + "A"], cn, REF[pre + "pm"], pn, len(pn))
    else:
        kf.UpdatePoints(REF[pre + "pm"], pn, len(pn))
    xg, Pg = kf.getState(), kf.getCov()
    assert rel(xg, REF[pre + "x_ref_upd"]) < TOL and rel(Pg, REF[pre + "P_ref_upd"]) < TOL
    assert np.array_equal(Pg, Pg.T)
    kf.close()


def test_augmentation_vs_reference_kalmanclass_fixtures():
    kf = ekf.KalmanFilter(capacity=64)
    kf.AddFirstStates(REF["aug_z19"])
    assert np.array_equal(kf.getState(), REF["aug_x_ref0"]) and np.array_equal(kf.getCov(), REF["aug_P_ref0"])
    kf.set_state(REF["aug_x_in"], REF["aug_P_in"], [12, 20], [])
    kf.AddNewCurve(REF["aug_z8"], REF["aug_A"])
    assert rel(kf.getState(), REF["aug_x_ref1"]) < TOL and rel(kf.getCov(), REF["aug_P_ref1"]) < TOL
    kf.AddNewPoints(REF["aug_pts"], 3)
    assert rel(kf.getState(), REF["aug_x_ref2"]) < TOL and rel(kf.getCov(), REF["aug_P_ref2"]) < TOL
    assert kf.curve_inds == REF["aug_ci_ref"].tolist() and kf.point_inds == REF["aug_pi_ref"].tolist()
    for t, ref in zip(REF["split_t"], REF["split_ref"]):
        A1, A2 = kf.GetSplitMatrices(float(t))
        assert np.array_equal(A1, ref[0]) and np.array_equal(A2, ref[1])
    kf.close()


@pytest.mark.parametrize("L,M_pts", [(1000, 0), (1000, 10)])
def test_update_at_1k_landmarks_vs_oracle(L, M_pts):
    """N = 3012 (1k point landmarks + pose; BASELINE configs[2]): n = 4 curves (M = 35), optionally 10 point
    measurements (M = 65); sequence predict -> update -> update against the dense CPU restatement."""
    x, G, ci, pi = ekf.make_synthetic_ekf(L, d=3, seed=5)
    N = x.size
    P = G @ G.T / 64.0 + 0.25 * np.eye(N)
    P = 0.5 * (P + P.T)
    kf = ekf.KalmanFilter(capacity=N)
    kf.set_state(x, None, ci, pi)
    kf.set_cov_lowrank(G, 1.0 / 64.0, 0.25)
    assert rel(kf.getCov(), P) < 1e-13
    rng = np.random.default_rng(1)
    A = np.stack([np.eye(4)] * 4)
    xo, Po = x.copy(), P.copy()
    for step in range(2):
        kf.PredictKF()
        xo, Po = oracle().ekf_predict(xo, Po)
        z = np.concatenate([rng.normal(0, 0.5, 32) + np.concatenate([xo[c:c + 8] for c in ci]) * 0.0, xo[2:5]])
        # measurement = prediction + noise (prediction computed by the oracle's H z-hat path is implicit)
        pts = rng.choice(len(pi), M_pts, replace=False) if M_pts else np.zeros(0, dtype=int)
        pm = rng.normal(0, 3.0, 3 * M_pts)
        kf.UpdateNCurvesAndPoints(z, 4, A, [0, 1, 2, 3], pm, pts, M_pts)
        xo, Po = oracle().ekf_update_curves_points(xo, Po, z, A, ci, pm, pi[pts] if M_pts else None)
        if step == 0:  # one update: the element-wise bar
            assert rel(kf.getState(), xo) < TOL and rel(kf.getCov(), Po) < TOL
    # two consecutive updates with wild innovations: the rounding of the first S^-1 (cond(S) ~ 5e4) is input to the
    # second, so the element-wise distance between two float64 implementations compounds; the max-norm bar holds
    xg, Pg = kf.getState(), kf.getCov()
    from parity import maxnorm_rel
    assert maxnorm_rel(xg, xo) < TOL and maxnorm_rel(Pg, Po) < TOL
    assert rel(xg, xo) < 1e-7 and rel(Pg, Po) < 1e-7
    assert np.array_equal(Pg, Pg.T)
    kf.close()


def test_update_properties_at_5k_landmarks():
    """N = 15012: too large for the dense CPU restatement inside a test; size-independent properties instead:
    the result is exactly symmetric, the trace never increases (K S K^T is PSD), the measured pose rows agree
    with an independent numpy evaluation of the sparse formulas in float64."""
    x, G, ci, pi = ekf.make_synthetic_ekf(5000, d=3, seed=6)
    N = x.size
    kf = ekf.KalmanFilter(capacity=N)
    kf.set_state(x, None, ci, pi)
    kf.set_cov_lowrank(G, 1.0 / 64.0, 0.25)
    d0 = np.array([kf.getCovBlock(i, i, 1, 1)[0, 0] for i in range(0, N, 997)])
    rows_before = kf.getCovBlock(0, 0, 12, N)
    rng = np.random.default_rng(2)
    z = np.concatenate([rng.normal(0, 1.0, 32), x[2:5]])
    kf.UpdateNCurvesAndPoints(z, 4, np.stack([np.eye(4)] * 4), [0, 1, 2, 3])
    kf.sync()
    d1 = np.array([kf.getCovBlock(i, i, 1, 1)[0, 0] for i in range(0, N, 997)])
    assert np.all(d1 <= d0 + 1e-12)
    blk = kf.getCovBlock(0, 0, 200, 200)
    assert np.array_equal(blk, blk.T)
    top = kf.getCovBlock(0, N - 300, 12, 300)
    left = kf.getCovBlock(N - 300, 0, 300, 12)
    assert np.array_equal(top, left.T)
    assert np.isfinite(rows_before).all() and np.isfinite(top).all()
    kf.close()


def _gate_inputs(L, nz, seed):
    x, G, ci, pi = ekf.make_synthetic_ekf(L, d=3, seed=seed)
    rng = np.random.default_rng(seed)
    N = x.size
    P = G @ G.T / 64.0 * 0.05 + 0.25 * np.eye(N)
    P = 0.5 * (P + P.T)
    pack = oracle().gate_pack(x, P, pi)
    pick = rng.integers(0, len(pi), nz)
    z = pack[pick, :3] + rng.normal(0, 5.0, (nz, 3))
    clutter = rng.random(nz) < 0.2
    z[clutter] = rng.uniform(-50, 50, (int(clutter.sum()), 3)) + rng.choice([-300.0, 300.0], (int(clutter.sum()), 3))
    return x, P, ci, pi, z


@pytest.mark.parametrize("L,nz", [(300, 257), (2000, 1000)])
def test_gate_bit_exact_indices(L, nz):
    x, P, ci, pi, z = _gate_inputs(L, nz, seed=7)
    kf = ekf.KalmanFilter(capacity=x.size)
    kf.set_state(x, P, ci, pi)
    idx, d2 = kf.gate(z, 11.345)
    io, do = oracle().gate(x, P, pi, z, 11.345, nthreads=8)
    assert np.array_equal(idx, io)
    assert np.array_equal(d2.view(np.int64), do.view(np.int64))
    assert (idx >= 0).any() and (idx < 0).any()
    # ties -> lowest index: duplicate measurements of exact predicted positions
    pack = oracle().gate_pack(x, P, pi)
    idx2 = kf.gate(pack[:50, :3], 11.345, want_d2=False)
    assert np.array_equal(idx2, oracle().gate(x, P, pi, pack[:50, :3], 11.345)[0])
    assert kf.gate(np.zeros((0, 3)), 11.345, want_d2=False).size == 0
    kf.close()


def test_augmentation_sequence_vs_golden():
    """AddFirstStates -> (state overwritten to the golden input) -> AddNewCurve -> AddNewPoints -> update, in place
    on the device, against the cv2 transcription of kalmanClass.cpp:275-666."""
    g = np.load(os.path.join(HERE, "golden", "ekf_augment.npz"))
    kf = ekf.KalmanFilter(capacity=64)
    kf.AddFirstStates(g["z19"])
    assert kf.N == 28 and kf.curve_inds == [12, 20] and kf.num_points == 0
    assert np.array_equal(kf.getState(), g["x0"]) and np.array_equal(kf.getCov(), g["P0"])
    kf.set_state(g["x_in"], g["P_in"], [12, 20], [])
    kf.AddNewCurve(g["z8"], g["A"])
    assert kf.N == 36 and kf.curve_inds == [12, 20, 28]
    assert rel(kf.getState(), g["x1"]) < TOL and rel(kf.getCov(), g["P1"]) < TOL
    kf.AddNewPoints(g["pts"], 3)
    assert kf.N == 45 and kf.point_inds == [36, 39, 42]
    x, P = kf.getState(), kf.getCov()
    assert rel(x, g["x2"]) < TOL and rel(P, g["P2"]) < TOL
    # the grown filter keeps working: predict + update of the new curve against the oracle on the same state
    z = np.concatenate([g["z8"] + 0.05, x[2:5]])
    kf.PredictKF()
    kf.UpdateNCurvesAndPoints(z, 1, np.eye(4)[None], [2])
    xo, Po = oracle().ekf_predict(x, P)
    xo, Po = oracle().ekf_update_curves_points(xo, Po, z, np.eye(4)[None], [28])
    assert rel(kf.getState(), xo) < TOL and rel(kf.getCov(), Po) < TOL
    with pytest.raises(Exception):
        for _ in range(10):
            kf.AddNewCurve(g["z8"], g["A"])  # capacity 64 is exhausted after three more curves
    kf.close()


def test_update_curve_landmarks_with_split_matrices():
    """8-state curve landmarks (N = 12 + 8*100) measured through de Casteljau split matrices (main.cpp:715-755 feeds
    A1/A2 of GetSplitMatrices to the update) against the dense CPU restatement."""
    x, G, ci, pi = ekf.make_synthetic_ekf(100, d=8, seed=8)
    N = x.size
    P = G @ G.T / 64.0 + 0.25 * np.eye(N)
    P = 0.5 * (P + P.T)
    kf = ekf.KalmanFilter(capacity=N)
    kf.set_state(x, P, ci, pi)
    A1, A2 = kf.GetSplitMatrices(0.35)
    A = np.stack([A1, A2, A2, A1])
    which = [3, 17, 42, 99]
    rng = np.random.default_rng(3)
    z = np.concatenate([rng.normal(0, 20.0, 32), x[2:5] + rng.normal(0, 0.05, 3)])
    kf.UpdateNCurvesAndPoints(z, 4, A, which)
    xo, Po = oracle().ekf_update_curves_points(x, P, z, A, ci[which])
    assert rel(kf.getState(), xo) < TOL and rel(kf.getCov(), Po) < TOL
    kf.close()


def test_failed_update_leaves_the_filter_untouched():
    """An innovation covariance that is not positive definite (here: an indefinite P) fails the update: the next
    sync reports it, and x and P are exactly what they were -- no NaN reaches the device-resident state (the
    reference's cvInvert(CV_SVD) would return a pseudo-inverse and carry on with a meaningless but finite state)."""
    g, ci, pi = load("ekf_curves4")
    N = g["P"].shape[0]
    Pbad = -10.0 * np.eye(N)
    kf = ekf.KalmanFilter(capacity=N)
    kf.set_state(g["x"], Pbad, ci, pi)
    curve_num = [int(np.where(ci == c)[0][0]) for c in g["curve_col"]]
    kf.UpdateNCurvesAndPoints(g["z"], len(curve_num), g["A"], curve_num)
    with pytest.raises(Exception):
        kf.sync()
    assert np.array_equal(kf.getState(), g["x"]) and np.array_equal(kf.getCov(), Pbad)
    # the handle keeps working afterwards
    kf.set_state(g["x"], g["P"], ci, pi)
    kf.UpdateNCurvesAndPoints(g["z"], len(curve_num), g["A"], curve_num)
    kf.sync()
    assert rel(kf.getState(), g["x_upd"]) < TOL and rel(kf.getCov(), g["P_upd"]) < TOL
    kf.close()


@pytest.mark.parametrize("M_pts", [0, 10, 15])
def test_tma_half_read_kernel_is_bit_identical_to_the_two_tile_kernel(monkeypatch, M_pts):
    """The TMA kernel reads one tile of every mirrored pair (12 N^2 bytes) because the tiles are exact mirrors after any
    update; its P must equal, bit for bit, what the cp.async / DMMA kernels that read both tiles (16 N^2 bytes) produce.
    N = 12 012: large enough for the persistent kernels (>= 2 pairs per SM), ragged last tile.  M = 35 runs the
    three-stage ring, M = 65 (4 curves + 10 points, BASELINE configs[2]) and M = 80 the two-stage ring."""
    x, G, ci, pi = ekf.make_synthetic_ekf(4000, d=3, seed=11)
    rng = np.random.default_rng(5)
    A = np.stack([np.eye(4)] * 4)
    z1 = np.concatenate([rng.normal(0, 1.0, 32), x[2:5]])
    z2 = np.concatenate([rng.normal(0, 1.0, 32), x[2:5]])
    pts = rng.choice(len(pi), M_pts, replace=False) if M_pts else None
    pm1, pm2 = rng.normal(0, 3.0, 3 * M_pts), rng.normal(0, 3.0, 3 * M_pts)
    res = {}
    for mode in ("tma", "two_tile"):
        if mode == "two_tile":
            monkeypatch.setenv("CPS_EKF_COV_NO_TMA", "1")
        kf = ekf.KalmanFilter(capacity=x.size)
        kf.set_state(x, None, ci, pi)
        kf.set_cov_lowrank(G, 1.0 / 64.0, 0.25)
        kf.PredictKF()
        kf.UpdateNCurvesAndPoints(z1, 4, A, [0, 1, 2, 3], pm1 if M_pts else None, pts, M_pts)
        kf.PredictKF()
        kf.UpdateNCurvesAndPoints(z2, 4, A, [0, 1, 2, 3], pm2 if M_pts else None, pts, M_pts)
        res[mode] = (kf.getState(), kf.getCov())
        kf.close()
    assert np.array_equal(res["tma"][0], res["two_tile"][0])
    assert np.array_equal(res["tma"][1].view(np.int64), res["two_tile"][1].view(np.int64))
    assert np.array_equal(res["tma"][1], res["tma"][1].T)


def test_fused_front_end_matches_the_five_kernel_path(monkeypatch):
    """build + S + S^-1 in one block and P H^T + gain in one pass against kernels 1-4 (Cholesky S^-1): same H, T, S;
    S^-1 by another elimination, so x and P agree to rounding."""
    g, ci, pi = load("ekf_curves2_points3")
    out = {}
    for mode in ("fused", "five"):
        if mode == "five":
            monkeypatch.setenv("CPS_EKF_NO_FUSED_FRONT", "1")
        kf = ekf.KalmanFilter(capacity=g["P"].shape[0])
        kf.set_state(g["x"], g["P"], ci, pi)
        curve_num = [int(np.where(ci == c)[0][0]) for c in g["curve_col"]]
        point_num = [int(np.where(pi == c)[0][0]) for c in g["point_col"]]
        kf.UpdateNCurvesAndPoints(g["z"], len(curve_num), g["A"], curve_num, g["point_meas"], point_num, len(point_num))
        out[mode] = (kf.getState(), kf.getCov(), kf.last_timing()["launches"])
        kf.close()
    assert out["fused"][2] == 3 and out["five"][2] == 5
    assert rel(out["fused"][0], out["five"][0]) < 1e-11 and rel(out["fused"][1], out["five"][1]) < 1e-11


def test_half_written_covariance_is_invisible(monkeypatch):
    """By default the streaming covariance kernel writes only tile (I,J), I <= J, of every mirrored pair (8 N^2 bytes per
    update; cps_ekf::lower_stale) and everything on the predict / update / gate / augmentation path reads the upper
    copy; the lower tiles are rebuilt when an entry point needs them.  Against CPS_EKF_EAGER_LOWER=1 (both tiles
    written by every update) every observable must be the same bits: state, covariance, blocks, gate, and a sequence
    that continues after the covariance was read back."""
    x, G, ci, pi = ekf.make_synthetic_ekf(3000, d=3, seed=21)
    rng = np.random.default_rng(9)
    A = np.stack([np.eye(4)] * 4)
    zs = [np.concatenate([rng.normal(0, 1.0, 32), x[2:5]]) for _ in range(3)]
    pts = rng.choice(len(pi), 5, replace=False)
    pm = rng.normal(0, 3.0, 15)
    gz = np.stack([x[pi[k]:pi[k] + 3] for k in rng.choice(len(pi), 64, replace=False)]) + rng.normal(0, 0.3, (64, 3))
    new_pts = rng.normal(0, 5.0, 6)
    out = {}
    for mode in ("lazy", "eager"):
        if mode == "eager":
            monkeypatch.setenv("CPS_EKF_EAGER_LOWER", "1")
        kf = ekf.KalmanFilter(capacity=x.size + 64)
        kf.set_state(x, None, ci, pi)
        kf.set_cov_lowrank(G, 1.0 / 64.0, 0.25)
        kf.PredictKF()
        kf.UpdateNCurvesAndPoints(zs[0], 4, A, [0, 1, 2, 3])
        kf.PredictKF()
        kf.UpdateNCurvesAndPoints(zs[1], 4, A, [0, 1, 2, 3], pm, pts, 5)  # M = 50: the two-stage ring
        r = {"upper_block": kf.getCovBlock(0, 64, 64, 200),    # inside the upper tiles: no rebuild needed
             "gate": kf.gate(gz, 11.345)}
        kf.AddNewPoints(new_pts, 2)                               # augmentation on a half-written P
        kf.PredictKF()
        kf.UpdatePoints(pm[:6], [len(pi), len(pi) + 1], 2)       # touches the new landmarks
        r["lower_block"] = kf.getCovBlock(kf.N - 100, 0, 100, 70)  # below the diagonal tiles: rebuilt on demand
        r["P1"] = kf.getCov()
        kf.PredictKF()
        kf.UpdateNCurvesAndPoints(zs[2], 4, A, [0, 1, 2, 3])     # the sequence goes on after the read-back
        r["x2"], r["P2"] = kf.getState(), kf.getCov()
        out[mode] = r
        kf.close()
    a, b = out["lazy"], out["eager"]
    for k in ("upper_block", "lower_block", "P1", "x2", "P2"):
        assert np.array_equal(a[k].view(np.int64), b[k].view(np.int64)), k
    assert np.array_equal(a["gate"][0], b["gate"][0]) and np.array_equal(a["gate"][1].view(np.int64), b["gate"][1].view(np.int64))
    assert np.array_equal(a["P2"], a["P2"].T) and np.array_equal(a["lower_block"], a["P1"][-100:, :70])
