#!/usr/bin/env python
"""Generates tests/golden/ref_cps.npz with the REFERENCE's own C++: kalmanClass.cpp and curveFittingClass.cpp compiled
UNCHANGED from /root/reference into oracle/_ref/libcps_ref.so (`make -C oracle ref`; OpenCV-1.x / GSL stand-ins under
oracle/shim/, harness oracle/ref_harness.cpp).  Every array named `*_ref*` below is an output of the reference's
code; the rest are the inputs it was given.  Runs only where /root/reference exists; the vectors are committed and
checked against the oracle restatements by tests/test_ref_pinning.py (CPU) and against the CUDA path by the -m gpu
tests.

    python tests/golden/gen_ref_golden.py
"""
import contextlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from curvepointslam_b200 import synth  # noqa: E402
from oracle_lib import build_oracle, refcps  # noqa: E402


@contextlib.contextmanager
def quiet():
    """The reference prints debug lines (kalmanClass.cpp:1041-1050); keep them out of the log."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        yield
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)


def ekf_case(seed, nc, npt, meas_curves, meas_points, split_t, rank=16):
    """Same recipe as gen_ekf_golden.make_case; layout: curves first, then points."""
    rng = np.random.default_rng(seed)
    N = 12 + 8 * nc + 3 * npt
    G = rng.standard_normal((N, rank))
    P = G @ G.T / rank + 0.25 * np.eye(N)
    P = 0.5 * (P + P.T)
    x = np.zeros(N)
    x[:12] = (0.3, -0.2, -1.0, 0.01, -0.02, 0.3, 1.0, 0.05, 0.0, 0.01, 0.0, 0.02)
    for c in range(nc):
        b = 12 + 8 * c
        x[b:b + 4] = np.linspace(2.0, 5.0, 4) + rng.normal(0, 0.1, 4)
        x[b + 4:b + 8] = (-1.2 if c % 2 == 0 else 1.2) + rng.normal(0, 0.05, 4)
    for p in range(npt):
        b = 12 + 8 * nc + 3 * p
        x[b:b + 3] = rng.uniform(-10, 10, 3)
    ci = 12 + 8 * np.arange(nc, dtype=np.int32)
    pi = 12 + 8 * nc + 3 * np.arange(npt, dtype=np.int32)
    return dict(x=x, P=P, ci=ci, pi=pi, cn=np.array(meas_curves, dtype=np.int32), pn=np.array(meas_points, dtype=np.int32),
                split_t=np.array(split_t, dtype=np.float64), rng=rng)


def main():
    build_oracle(with_ref=True)
    r = refcps()
    assert r is not None, "oracle/_ref/libcps_ref.so missing: needs /root/reference"
    out = {}

    # ---- EKF: predict, curve(+point) update, point update, on four maps ------------------------------------------
    cases = {
        "c4": ekf_case(1, 6, 0, [0, 1, 4, 5], [], [0.3, 0.3, 0.6, 0.6]),
        "c2p3": ekf_case(2, 3, 5, [2, 0], [4, 0, 2], [0.5, 0.25]),
        "p4": ekf_case(3, 2, 6, [], [1, 5, 3, 0], []),
        "c4p10": ekf_case(4, 8, 60, [7, 2, 3, 0], [5, 11, 17, 23, 29, 35, 41, 47, 53, 59], [0.2, 0.4, 0.6, 0.8], rank=32),
    }
    for name, c in cases.items():
        kf = r.kf()
        kf.set(c["x"], c["P"], c["ci"], c["pi"])
        A = np.zeros((len(c["cn"]), 4, 4))
        for k, t in enumerate(c["split_t"]):
            A[k] = kf.split_matrices(t)[k % 2]
        n = len(c["cn"])
        z = np.zeros(8 * n + 3)
        for k in range(n):
            z[8 * k:8 * k + 8] = kf.predicted_measurement(A[k], int(c["cn"][k])) + c["rng"].normal(0, 0.5, 8)
        z[8 * n:] = c["x"][2:5] + c["rng"].normal(0, 0.1, 3)
        pm = np.zeros(3 * len(c["pn"]))
        R_be = r.generate_rbe(c["x"][3], c["x"][4], c["x"][5])
        for k, j in enumerate(c["pn"]):
            col = c["pi"][j]
            pm[3 * k:3 * k + 3] = R_be @ (c["x"][col:col + 3] - c["x"][:3]) + c["rng"].normal(0, 2.0, 3)
        pre = f"ekf_{name}_"
        out.update({pre + "x": c["x"], pre + "P": c["P"], pre + "ci": c["ci"], pre + "pi": c["pi"], pre + "cn": c["cn"],
                    pre + "pn": c["pn"], pre + "A": A, pre + "z": z, pre + "pm": pm})
        with quiet():
            kf.predict()
        out[pre + "x_ref_pred"], out[pre + "P_ref_pred"], _, _ = kf.get()
        kf.set(c["x"], c["P"], c["ci"], c["pi"])
        with quiet():
            if n:
                kf.update_curves_points(z, A, c["cn"], pm, c["pn"])
            else:
                kf.update_points(pm, c["pn"])
        out[pre + "x_ref_upd"], out[pre + "P_ref_upd"], _, _ = kf.get()
        if n:
            out[pre + "zhat_ref"] = np.stack([kf.predicted_measurement(A[k], int(c["cn"][k])) for k in range(n)])
        kf.close()

    # ---- EKF augmentation chain: AddFirstStates -> PredictKF -> AddNewCurve -> AddNewPoints --------------------
    rng = np.random.default_rng(11)
    z19 = np.concatenate([np.linspace(2.0, 5.0, 4), -1.2 + rng.normal(0, 0.05, 4), np.linspace(2.1, 5.2, 4),
                          1.2 + rng.normal(0, 0.05, 4), [-1.0, 0.01, -0.02]])
    kf = r.kf()
    with quiet():
        kf.add_first_states(z19)
        x0, P0, _, _ = kf.get()
        kf.predict()
        x_in, P_in, _, _ = kf.get()
        x_in[5] = 0.2
        kf.set(x_in, P_in, [12, 20], [])
        A1, A2 = kf.split_matrices(0.4)
        z8 = np.concatenate([np.linspace(3.0, 6.0, 4), -1.1 + rng.normal(0, 0.05, 4)])
        kf.add_new_curve(z8, A2)
        x1, P1, ci1, _ = kf.get()
        pts = rng.uniform(-5, 5, 9)
        kf.add_new_points(pts)
        x2, P2, ci2, pi2 = kf.get()
    out.update(aug_z19=z19, aug_x_ref0=x0, aug_P_ref0=P0, aug_x_in=x_in, aug_P_in=P_in, aug_z8=z8, aug_A=A2,
               aug_x_ref1=x1, aug_P_ref1=P1, aug_pts=pts, aug_x_ref2=x2, aug_P_ref2=P2, aug_ci_ref=ci2, aug_pi_ref=pi2)
    ts = np.array([0.1, 0.37, 0.5, 0.9])
    out["split_t"] = ts
    out["split_ref"] = np.stack([np.stack(kf.split_matrices(t)) for t in ts])
    kf.close()
    ang = np.array([[0.01, -0.02, 0.3], [0.4, 0.2, -1.1], [-0.7, 1.0, 2.5]])
    out["rbe_angles"] = ang
    out["rbe_ref"] = np.stack([r.generate_rbe(*a) for a in ang])
    out["rbe_derivs_ref"] = np.stack([np.stack(r.rbe_derivs(*a)) for a in ang])

    # ---- curve model: modelfunc / cp_earth2image on random parameters ---------------------------------------------
    b = synth.make_stereo19_batch(12, seed=7, round_pixels=True,
                                  counts=np.random.default_rng(3).integers(12, 90, size=(12, 4)).astype(np.int32))
    t_all = synth.t_from_rows(b)
    rng = np.random.default_rng(0)
    mp, mhx = [], []
    for f in range(12):
        p = b["p0"][f] + rng.normal(0, 0.01, 19)
        t = t_all[b["off"][f] // 2:b["off"][f + 1] // 2]
        mp.append(p)
        mhx.append(r.modelfunc(p, t, b["counts"][f]))
    out.update(model_meas=b["meas"], model_off=b["off"], model_counts=b["counts"], model_t=t_all, model_p=np.stack(mp),
               model_hx_ref=np.concatenate(mhx))
    cl, cr = zip(*[r.cp_earth2image(p[:8], [0.0, p[16], p[17]], [0.0, 0.0, p[18]]) for p in mp])
    out.update(model_cleft_ref=np.stack(cl), model_cright_ref=np.stack(cr))

    # ---- the Bernstein evaluation alone = the m = 8 per-contour model (closest_bezier_pt(cp, t, pt, out), :1057-1073)
    rng = np.random.default_rng(4)
    bcp = rng.uniform(20, 300, (50, 8))
    bt = rng.uniform(0, 1, (50, 32))
    bt[:, 0], bt[:, -1] = 0.0, 1.0
    out.update(bez_cp=bcp, bez_t=bt, bez_xy_ref=np.stack([np.stack([r.bezier_eval(bcp[i], t) for t in bt[i]]) for i in range(50)]))

    # ---- fit_curve end to end (the reference's live dlevmar_dif call) ------------------------------------------
    fb = synth.make_stereo19_batch(16, K=64, seed=21, round_pixels=True)
    fp0, ft, fplm, finfo, fret, fpar, ferr = [], [], [], [], [], [], []
    for f in range(16):
        meas = fb["meas"][fb["off"][f]:fb["off"][f + 1]]
        p, err = r.fit_curve(meas, fb["counts"][f])
        h = r.last_fit()
        assert np.array_equal(h["x"], meas) and np.array_equal(h["counts"], fb["counts"][f])
        fp0.append(h["p0"]); ft.append(h["t"]); fplm.append(h["p_lm"]); finfo.append(h["info"]); fret.append(h["ret"])
        fpar.append(p); ferr.append(err)
    out.update(fit_meas=fb["meas"], fit_off=fb["off"], fit_counts=fb["counts"], fit_p0_ref=np.stack(fp0),
               fit_t_ref=np.concatenate(ft), fit_plm_ref=np.stack(fplm), fit_info_ref=np.stack(finfo),
               fit_ret_ref=np.array(fret, dtype=np.int32), fit_params_ref=np.stack(fpar), fit_err_ref=np.array(ferr),
               fit_opts_ref=h["opts"], fit_itmax_ref=np.array(h["itmax"]))

    # ---- cleanup_and_group_edges ---------------------------------------------------------------------------------
    F = 160
    pts_e, so, tr = synth.make_edge_frames(F, seed=11, consistent=False)
    for reset in (0, 1):
        ok, cnt, ms = [], [], []
        for f in range(F):
            o_, c_, m_, _, _ = r.cleanup_edges(pts_e, so[4 * f:4 * f + 5], tr[f], bool(reset))
            ok.append(o_); cnt.append(c_); ms.append(m_)
        out[f"edges_ok_ref{reset}"] = np.array(ok, dtype=np.int32)
        out[f"edges_counts_ref{reset}"] = np.stack(cnt)
        out[f"edges_meas_ref{reset}"] = np.concatenate(ms)
    out.update(edges_pts=pts_e, edges_seg_off=so, edges_tracked=tr)

    # ---- closest point on a cubic (control2coeffs + quintic; the root finder is the shared GSL restatement) ------
    rng = np.random.default_rng(5)
    Q = 400
    cps = np.array([100, 200, 120, 150, 160, 100, 200, 60], dtype=float) + rng.normal(0, 15, (Q, 8))
    qpts = rng.uniform(40, 220, (Q, 2))
    coef = np.stack([r.control2coeffs(c) for c in cps])
    res = [r.closest(coef[q], qpts[q]) for q in range(Q)]
    out.update(closest_cp=cps, closest_pt=qpts, closest_coef_ref=coef, closest_near_ref=np.stack([a for a, _, _ in res]),
               closest_dist_ref=np.array([d for _, d, _ in res]), closest_t_ref=np.array([t for _, _, t in res]))

    path = os.path.join(HERE, "ref_cps.npz")
    np.savez_compressed(path, **out)
    print("written", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
