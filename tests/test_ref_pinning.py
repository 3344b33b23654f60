"""CPU tests that pin the non-levmar oracle restatements (orc_ekf.c, orc_model.c, orc_closest.c) to the REFERENCE'S OWN
C++: /root/reference/kalmanClass.cpp and curveFittingClass.cpp compiled unchanged into oracle/_ref/libcps_ref.so
(oracle/Makefile `ref`, stand-ins for the OpenCV-1.x / GSL calls under oracle/shim/).

Two layers:
  * test_fixture_*: the oracle against tests/golden/ref_cps.npz, outputs of the reference build written by
    tests/golden/gen_ref_golden.py and committed -- these run everywhere;
  * test_live_*: where the reference build is present (this container; it also travels to the GPU box), the
    reference is re-run and must reproduce the committed fixture bit for bit, and is compared with the oracle on
    fresh random inputs.

Tolerances.  Integer / index work and everything without a matrix inverse or SVD: bit-exact.  With an inverse
(cvInvert(CV_SVD) is a Jacobi SVD pseudo-inverse in the stand-in, a Jacobi SVD of its own in the oracle): element-wise
|a - b| <= 1e-9 * max(|b|, 1e-2 * max|b|) (the north star's 1e-9 relative, per element; tests/parity.py says why 1e-2), observed ~1e-13 max-norm.
"""
import ctypes as C
import os

import numpy as np
import pytest

from curvepointslam_b200 import synth
from oracle_lib import DIF, MATH_DET, MATH_LIBM, STEREO19, oracle, refcps
from parity import elementwise_rel

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ref_cps.npz"))
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
need_ref = pytest.mark.skipif(refcps() is None, reason="oracle/_ref/libcps_ref.so not built (needs /root/reference)")
EKF_CASES = ["c4", "c2p3", "p4", "c4p10"]


def elementwise_ok(a, b, rtol=1e-9):
    """tests/parity.py: |a - b| <= rtol * max(|b|, 1e-2 * max|b|), relative per element"""
    return elementwise_rel(a, b) <= rtol


def maxrel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


class FitData(C.Structure):
    _fields_ = [("t", dp), ("count", C.c_int * 4), ("math_mode", C.c_int)]


def oracle_model(p, t, counts, mode):
    t = np.ascontiguousarray(t)
    fd = FitData(t.ctypes.data_as(dp), (C.c_int * 4)(*[int(c) for c in counts]), mode)
    n = 2 * int(np.sum(counts))
    hx, pp = np.zeros(n), np.array(p, dtype=np.float64)
    oracle().lib.orc_stereo19_func(pp.ctypes.data_as(dp), hx.ctypes.data_as(dp), 19, n, C.byref(fd))
    return hx


def oracle_ekf_update(pre):
    o = oracle()
    x, P = G[pre + "x"], G[pre + "P"]
    cc = G[pre + "ci"][G[pre + "cn"]] if len(G[pre + "cn"]) else np.zeros(0, dtype=np.int32)
    pc = G[pre + "pi"][G[pre + "pn"]] if len(G[pre + "pn"]) else np.zeros(0, dtype=np.int32)
    if len(cc):
        return o.ekf_update_curves_points(x, P, G[pre + "z"], G[pre + "A"], cc, G[pre + "pm"], pc)
    return o.ekf_update_points(x, P, G[pre + "pm"], pc)


# ---------------------------------------------------------------------------------------------------- fixtures
@pytest.mark.parametrize("name", EKF_CASES)
def test_fixture_ekf_predict_is_bit_identical_to_the_reference(name):
    """PredictKF (kalmanClass.cpp:161-272): 12x12 / 12xN products, no inverse -> identical doubles."""
    pre = f"ekf_{name}_"
    x, P = oracle().ekf_predict(G[pre + "x"], G[pre + "P"])
    assert np.array_equal(x, G[pre + "x_ref_pred"]) and np.array_equal(P, G[pre + "P_ref_pred"])


@pytest.mark.parametrize("name", EKF_CASES)
def test_fixture_ekf_update_matches_the_reference(name):
    """UpdateNCurvesAndPoints / UpdatePoints (kalmanClass.cpp:669-945, 948-1126) on the reference's own build."""
    pre = f"ekf_{name}_"
    x, P = oracle_ekf_update(pre)
    assert elementwise_ok(x, G[pre + "x_ref_upd"]) and elementwise_ok(P, G[pre + "P_ref_upd"])
    assert maxrel(x, G[pre + "x_ref_upd"]) < 1e-12 and maxrel(P, G[pre + "P_ref_upd"]) < 1e-12
    assert np.array_equal(G[pre + "P_ref_upd"], G[pre + "P_ref_upd"].T)  # the reference's P is exactly symmetric too


def test_fixture_predicted_measurement_and_split_matrices():
    L = oracle().lib
    L.orc_ekf_predicted_curve.argtypes = [dp, dp, C.c_int, dp]
    for name in ("c4", "c2p3", "c4p10"):
        pre = f"ekf_{name}_"
        # zhat was recorded after the update: evaluate the restatement on the reference's updated state
        x = np.ascontiguousarray(G[pre + "x_ref_upd"])
        for k, cn in enumerate(G[pre + "cn"]):
            zh = np.zeros(8)
            A = np.ascontiguousarray(G[pre + "A"][k])
            L.orc_ekf_predicted_curve(x.ctypes.data_as(dp), A.ctypes.data_as(dp), int(G[pre + "ci"][cn]), zh.ctypes.data_as(dp))
            assert np.array_equal(zh, G[pre + "zhat_ref"][k])
    for t, ref in zip(G["split_t"], G["split_ref"]):
        A1, A2 = np.zeros(16), np.zeros(16)
        L.orc_ekf_split_matrices(float(t), A1.ctypes.data_as(dp), A2.ctypes.data_as(dp))
        assert np.array_equal(A1.reshape(4, 4), ref[0]) and np.array_equal(A2.reshape(4, 4), ref[1])


def test_fixture_rotation_and_its_derivative_tables_keep_the_reference_quirks():
    """generate_Rbe (common.h:174-186, sign of the (2,0)/(0,2) term) and get_Rbe_derivs (kalmanClass.cpp:1275-1298:
    R_be_theta(1,1) written twice, (2,1) never set)."""
    L = oracle().lib
    L.orc_ekf_rbe.argtypes = [C.c_double, C.c_double, C.c_double, dp]
    L.orc_ekf_rbe_derivs.argtypes = [C.c_double, C.c_double, C.c_double, dp, dp, dp]
    for a, R_ref, D_ref in zip(G["rbe_angles"], G["rbe_ref"], G["rbe_derivs_ref"]):
        R = np.zeros(9)
        L.orc_ekf_rbe(a[0], a[1], a[2], R.ctypes.data_as(dp))
        assert np.array_equal(R.reshape(3, 3), R_ref)
        d = [np.zeros(9) for _ in range(3)]
        L.orc_ekf_rbe_derivs(a[0], a[1], a[2], *[v.ctypes.data_as(dp) for v in d])
        for k in range(3):
            assert np.array_equal(d[k].reshape(3, 3), D_ref[k])
        assert D_ref[1][2, 1] == 0.0  # the entry the reference never sets


def test_fixture_augmentation_chain_matches_the_reference():
    """AddFirstStates -> AddNewCurve -> AddNewPoints (kalmanClass.cpp:275-666)."""
    L = oracle().lib
    L.orc_ekf_add_new_curve.restype = C.c_int
    L.orc_ekf_add_new_points.restype = C.c_int
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(dp)
    ld = 64
    x, P = np.zeros(ld), np.zeros((ld, ld))
    L.orc_ekf_add_first_states(x.ctypes.data_as(dp), P.ctypes.data_as(dp), ld, c(G["aug_z19"]))
    assert np.array_equal(x[:28], G["aug_x_ref0"]) and np.array_equal(P[:28, :28], G["aug_P_ref0"])
    x[:28], P[:28, :28] = G["aug_x_in"], G["aug_P_in"]
    N = L.orc_ekf_add_new_curve(x.ctypes.data_as(dp), P.ctypes.data_as(dp), ld, 28, c(G["aug_z8"]), c(G["aug_A"]))
    assert N == 36 and elementwise_ok(x[:N], G["aug_x_ref1"]) and elementwise_ok(P[:N, :N], G["aug_P_ref1"])
    N = L.orc_ekf_add_new_points(x.ctypes.data_as(dp), P.ctypes.data_as(dp), ld, N, c(G["aug_pts"]), 3)
    assert N == 45 and elementwise_ok(x[:N], G["aug_x_ref2"]) and elementwise_ok(P[:N, :N], G["aug_P_ref2"])
    assert G["aug_ci_ref"].tolist() == [12, 20, 28] and G["aug_pi_ref"].tolist() == [36, 39, 42]
    assert np.all(G["aug_P_ref2"][36:39, 39:45] == 0.0)  # the reference gives jointly added points no cross terms


def test_fixture_modelfunc_libm_mode_is_bit_identical_to_the_reference():
    """modelfunc + cp_earth2image + earth2body + Bernstein evaluation (curveFittingClass.cpp:82-119, 804-847,
    1057-1073): the oracle's LIBM mode reproduces every double; the DET mode (what the CUDA kernels compute: fixed
    sin/cos sequence, pow as multiplies) stays within a few ulp of it."""
    off, counts = G["model_off"], G["model_counts"]
    for f in range(len(counts)):
        t = G["model_t"][off[f] // 2:off[f + 1] // 2]
        ref = G["model_hx_ref"][off[f]:off[f + 1]]
        assert np.array_equal(oracle_model(G["model_p"][f], t, counts[f], MATH_LIBM), ref)
        det = oracle_model(G["model_p"][f], t, counts[f], MATH_DET)
        assert np.abs(det - ref).max() <= 4 * np.spacing(np.abs(ref).max())


def test_fixture_bernstein_model_m8_is_bit_identical_to_the_reference():
    """closest_bezier_pt(cp, t, pt, out) (curveFittingClass.cpp:1057-1073) = the m = 8 per-contour model."""
    for i in range(len(G["bez_cp"])):
        t = np.ascontiguousarray(G["bez_t"][i])
        fd = FitData(t.ctypes.data_as(dp), (C.c_int * 4)(t.size, 0, 0, 0), MATH_LIBM)
        hx, p = np.zeros(2 * t.size), np.ascontiguousarray(G["bez_cp"][i])
        oracle().lib.orc_image8_func(p.ctypes.data_as(dp), hx.ctypes.data_as(dp), 8, 2 * t.size, C.byref(fd))
        assert np.array_equal(hx.reshape(-1, 2), G["bez_xy_ref"][i])
        fd.math_mode = MATH_DET
        oracle().lib.orc_image8_func(p.ctypes.data_as(dp), hx.ctypes.data_as(dp), 8, 2 * t.size, C.byref(fd))
        assert np.abs(hx.reshape(-1, 2) - G["bez_xy_ref"][i]).max() <= 4 * np.spacing(np.abs(G["bez_xy_ref"][i]).max())


def test_fixture_fit_curve_is_bit_identical_to_the_reference():
    """CurveFittingClass::fit_curve end to end (curveFittingClass.cpp:408-724): packing, t, initial guess, the live
    dlevmar_dif call, post-normalisation.  From the reference's own initial guess the oracle (LIBM mode, dif) returns
    the reference's parameters, info[0..9] and return code bit for bit; its own initial guess differs from the
    reference's only by the rounding of two different SVD routines (the reference's cvSVD is external to it)."""
    o = oracle()
    off, counts = G["fit_off"], G["fit_counts"]
    B = len(counts)
    t_all = synth.t_from_rows(dict(meas=G["fit_meas"], off=off, counts=counts))
    assert np.array_equal(t_all, G["fit_t_ref"])  # t-parameterisation :451-511
    res = o.fit_batch(STEREO19, DIF, G["fit_p0_ref"], G["fit_meas"], off, counts, G["fit_opts_ref"],
                      itmax=int(G["fit_itmax_ref"]), math_mode=MATH_LIBM, nthreads=4)
    assert np.array_equal(res["p"], G["fit_plm_ref"])
    assert np.array_equal(res["info"], G["fit_info_ref"]) and np.array_equal(res["ret"], G["fit_ret_ref"])
    assert np.array_equal(res["info"][:, 1], G["fit_err_ref"])  # fitting_error = info[1]
    post = res["p"].copy()
    for f in range(B):
        o.lib.orc_stereo19_postfit(post[f].ctypes.data_as(dp))
    assert np.array_equal(post, G["fit_params_ref"])
    assert G["fit_opts_ref"].tolist() == [1e-10, 1e-10, 1e-10, 1e-20, 1e-6 * 1e1] and int(G["fit_itmax_ref"]) == 1000
    for f in range(B):  # initial guess :516-651
        m = np.ascontiguousarray(G["fit_meas"][off[f]:off[f + 1]])
        c = np.ascontiguousarray(counts[f], dtype=np.int32)
        p0 = np.zeros(19)
        o.lib.orc_stereo19_init_guess(m.ctypes.data_as(dp), c.ctypes.data_as(ip), p0.ctypes.data_as(dp), MATH_LIBM, 1)
        assert maxrel(p0, G["fit_p0_ref"][f]) < 1e-13


@pytest.mark.parametrize("reset", [0, 1])
def test_fixture_cleanup_and_group_edges_is_identical_to_the_reference(reset):
    """cleanup_and_group_edges (curveFittingClass.cpp:1207-1418): same verdict, same kept points, every frame."""
    r = oracle().cleanup_edges(G["edges_pts"], G["edges_seg_off"], G["edges_tracked"], reset=bool(reset))
    ok = G[f"edges_ok_ref{reset}"]
    assert np.array_equal(r["valid"], (ok == 1).astype(np.int32)) and 20 < ok.sum() < len(ok)
    assert np.array_equal(r["counts"][ok == 1], G[f"edges_counts_ref{reset}"][ok == 1])
    assert np.array_equal(r["meas"], G[f"edges_meas_ref{reset}"])


def test_fixture_closest_point_matches_the_reference():
    """control2coeffs + closest_bezier_pt(p, pt, ...) (curveFittingClass.cpp:936-1043): same root selected, same t;
    the distance differs by rounding only (the reference squares with pow(x, 2.0) and takes pow(d, 0.5))."""
    o = oracle()
    for q in range(len(G["closest_cp"])):
        p = np.zeros(8)
        cp = np.ascontiguousarray(G["closest_cp"][q])
        o.lib.orc_control2coeffs(cp.ctypes.data_as(dp), p.ctypes.data_as(dp))
        assert np.array_equal(p, G["closest_coef_ref"][q])
    near, dist, t, _ = o.closest(G["closest_cp"], G["closest_pt"])
    assert np.array_equal(t, G["closest_t_ref"])
    assert np.abs(near - G["closest_near_ref"]).max() <= 2e-13
    assert np.abs(dist - G["closest_dist_ref"]).max() <= 1e-13 * G["closest_dist_ref"].max()


# ---------------------------------------------------------------------------------------------------- live
@need_ref
def test_live_reference_reproduces_the_committed_fixture():
    r = refcps()
    off, counts = G["model_off"], G["model_counts"]
    for f in range(len(counts)):
        t = G["model_t"][off[f] // 2:off[f + 1] // 2]
        assert np.array_equal(r.modelfunc(G["model_p"][f], t, counts[f]), G["model_hx_ref"][off[f]:off[f + 1]])
    off, counts = G["fit_off"], G["fit_counts"]
    for f in (0, 5, 11):
        p, err = r.fit_curve(G["fit_meas"][off[f]:off[f + 1]], counts[f])
        assert np.array_equal(p, G["fit_params_ref"][f]) and err == G["fit_err_ref"][f]
    pre = "ekf_c4p10_"
    kf = r.kf()
    kf.set(G[pre + "x"], G[pre + "P"], G[pre + "ci"], G[pre + "pi"])
    kf.update_curves_points(G[pre + "z"], G[pre + "A"], G[pre + "cn"], G[pre + "pm"], G[pre + "pn"])
    x, P, _, _ = kf.get()
    kf.close()
    assert np.array_equal(x, G[pre + "x_ref_upd"]) and np.array_equal(P, G[pre + "P_ref_upd"])


@need_ref
def test_live_modelfunc_random_parameters_bit_identical():
    r = refcps()
    rng = np.random.default_rng(42)
    b = synth.make_stereo19_batch(6, K=48, seed=9, round_pixels=True)
    t_all = synth.t_from_rows(b)
    for f in range(6):
        t = t_all[b["off"][f] // 2:b["off"][f + 1] // 2]
        for _ in range(40):
            p = b["p0"][f] + rng.normal(0, 0.02, 19)
            assert np.array_equal(oracle_model(p, t, b["counts"][f], MATH_LIBM), r.modelfunc(p, t, b["counts"][f]))


@need_ref
def test_live_fit_curve_ragged_frames_bit_identical():
    r, o = refcps(), oracle()
    counts = np.random.default_rng(8).integers(10, 101, size=(10, 4)).astype(np.int32)
    b = synth.make_stereo19_batch(10, seed=33, counts=counts, round_pixels=True)
    for f in range(10):
        meas = b["meas"][b["off"][f]:b["off"][f + 1]]
        p_ref, err = r.fit_curve(meas, counts[f])
        h = r.last_fit()
        res = o.fit_batch(STEREO19, DIF, h["p0"][None], meas, np.array([0, meas.size]), counts[f][None], h["opts"],
                          itmax=h["itmax"], math_mode=MATH_LIBM)
        assert np.array_equal(res["p"][0], h["p_lm"]) and np.array_equal(res["info"][0], h["info"])
        assert res["ret"][0] == h["ret"]
        p = res["p"][0].copy()
        o.lib.orc_stereo19_postfit(p.ctypes.data_as(dp))
        assert np.array_equal(p, p_ref)


@need_ref
def test_live_ekf_update_at_a_thousand_states():
    """A denser map than the fixtures hold: N = 12 + 8*20 + 3*300 = 1072, 4 curves + 10 points (M = 65)."""
    r, o = refcps(), oracle()
    rng = np.random.default_rng(77)
    nc, npt = 20, 300
    N = 12 + 8 * nc + 3 * npt
    Gm = rng.standard_normal((N, 64))
    P = Gm @ Gm.T / 64 + 0.25 * np.eye(N)
    P = 0.5 * (P + P.T)
    x = np.zeros(N)
    x[:12] = (0.3, -0.2, -1.0, 0.01, -0.02, 0.3, 1.0, 0.05, 0.0, 0.01, 0.0, 0.02)
    x[12:] = rng.uniform(-20, 20, N - 12)
    ci = 12 + 8 * np.arange(nc, dtype=np.int32)
    pi = 12 + 8 * nc + 3 * np.arange(npt, dtype=np.int32)
    cn = np.array([3, 17, 0, 9], dtype=np.int32)
    pn = rng.choice(npt, 10, replace=False).astype(np.int32)
    kf = r.kf()
    kf.set(x, P, ci, pi)
    A = np.stack([kf.split_matrices(0.3 + 0.1 * k)[k % 2] for k in range(4)])
    z = np.concatenate([np.concatenate([kf.predicted_measurement(A[k], int(cn[k])) for k in range(4)]) +
                        rng.normal(0, 0.5, 32), x[2:5] + rng.normal(0, 0.1, 3)])
    R_be = r.generate_rbe(x[3], x[4], x[5])
    pm = np.concatenate([R_be @ (x[pi[j]:pi[j] + 3] - x[:3]) + rng.normal(0, 2.0, 3) for j in pn])
    kf.predict()
    kf.update_curves_points(z, A, cn, pm, pn)
    xr, Pr, _, _ = kf.get()
    kf.close()
    xo, Po = o.ekf_predict(x, P)
    xo, Po = o.ekf_update_curves_points(xo, Po, z, A, ci[cn], pm, pi[pn])
    assert elementwise_ok(xo, xr) and elementwise_ok(Po, Pr)
    assert maxrel(xo, xr) < 1e-12 and maxrel(Po, Pr) < 1e-12
