"""ctypes bindings of the CPU oracle (oracle/liborc.so) and, when present, the reference levmar built
unchanged into oracle/_ref/.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_DIR = os.path.join(ROOT, "oracle")

STEREO19, IMAGE8 = 19, 8
DER, DIF = 0, 1
MATH_DET, MATH_LIBM = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)
FUNC_T = C.CFUNCTYPE(None, _dp, _dp, C.c_int, C.c_int, C.c_void_p)


def build_oracle(with_ref: bool = True) -> None:
    subprocess.run(["make", "-s", "-C", ORC_DIR], check=True)
    if with_ref and os.path.isdir("/root/reference/levmar-2.5"):
        subprocess.run(["make", "-s", "-C", ORC_DIR, "ref"], check=True)


def _ptr(a, typ):
    if a is None:
        return None
    return a.ctypes.data_as(typ)


_FIT_ARGS = [C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _lp, _ip, C.c_int, _dp, _dp, _ip, _ip, C.c_int]


class Oracle:
    def __init__(self):
        path = os.path.join(ORC_DIR, "liborc.so")
        if not os.path.exists(path):
            build_oracle(with_ref=True)
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L = self.lib
        L.orc_fit_batch.argtypes = _FIT_ARGS
        L.orc_fit_batch.restype = C.c_int
        L.orc_lu_solve.argtypes = [_dp, _dp, _dp, C.c_int]
        L.orc_lu_solve.restype = C.c_int
        L.orc_residual_sqnorm.argtypes = [_dp, _dp, _dp, C.c_int]
        L.orc_residual_sqnorm.restype = C.c_double
        L.orc_ata_blocked.argtypes = [_dp, _dp, C.c_int, C.c_int]
        L.orc_kat_problem.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _ip, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.orc_kat_problem.restype = C.c_int
        L.orc_lm_der.argtypes = [C.c_void_p, C.c_void_p, _dp, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_void_p, C.c_void_p]
        L.orc_lm_der.restype = C.c_int
        L.orc_lm_dif.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_void_p, C.c_void_p]
        L.orc_lm_dif.restype = C.c_int
        L.orc_stereo19_postfit.argtypes = [_dp]
        L.orc_closest_bezier_pt.argtypes = [_dp, _dp, _dp, _dp]
        L.orc_closest_bezier_pt.restype = C.c_int
        L.orc_control2coeffs.argtypes = [_dp, _dp]
        L.orc_poly_roots.argtypes = [_dp, C.c_int, _dp, _dp]
        L.orc_poly_roots.restype = C.c_int
        L.orc_cleanup_edges.argtypes = [_ip, _lp, C.c_int, _dp, _ip]
        L.orc_cleanup_edges.restype = C.c_int
        L.orc_ekf_predict.argtypes = [_dp, _dp, C.c_int]
        L.orc_ekf_update_curves_points.argtypes = [_dp, _dp, C.c_int, _dp, C.c_int, _dp, _ip, _dp, _ip, C.c_int,
                                                   _dp, _dp, _dp]
        L.orc_ekf_update_curves_points.restype = C.c_int
        L.orc_ekf_update_points.argtypes = [_dp, _dp, C.c_int, _dp, _ip, C.c_int]
        L.orc_ekf_update_points.restype = C.c_int
        L.orc_ekf_split_matrices.argtypes = [C.c_double, _dp, _dp]
        L.orc_pinv_svd.argtypes = [_dp, _dp, C.c_int]
        L.orc_gate_mahalanobis.argtypes = [_dp, _dp, C.c_int, _ip, C.c_int, _dp, C.c_int, C.c_double, _ip, _dp,
                                           C.c_int]
        L.orc_gate_mahalanobis.restype = C.c_int
        L.orc_gate_pack.argtypes = [_dp, _dp, C.c_int, _ip, C.c_int, _dp]
        L.orc_gate_pack_pieces.argtypes = [_dp, _dp, _dp, _dp, C.c_int, _ip, C.c_int, _dp]
        L.orc_gate_mahalanobis_pieces.argtypes = [_dp, _dp, _dp, _dp, C.c_int, _ip, C.c_int, _dp, C.c_int, C.c_double,
                                                  _ip, _dp, C.c_int]
        L.orc_gate_mahalanobis_pieces.restype = C.c_int
        self.ref = None
        self.levmar = None
        rpath = os.path.join(ORC_DIR, "_ref", "liborc_refshim.so")
        lpath = os.path.join(ORC_DIR, "_ref", "liblevmar_ref.so")
        if os.path.exists(rpath) and os.path.exists(lpath):
            self.levmar = C.CDLL(lpath, mode=C.RTLD_GLOBAL)
            self.ref = C.CDLL(rpath)
            self.ref.ref_fit_batch.argtypes = _FIT_ARGS
            self.ref.ref_fit_batch.restype = C.c_int
            lv = self.levmar
            lv.dlevmar_der.argtypes = [C.c_void_p, C.c_void_p, _dp, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_void_p]
            lv.dlevmar_der.restype = C.c_int
            lv.dlevmar_dif.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_void_p]
            lv.dlevmar_dif.restype = C.c_int
            lv.dAx_eq_b_LU_noLapack.argtypes = [_dp, _dp, _dp, C.c_int]
            lv.dAx_eq_b_LU_noLapack.restype = C.c_int
            lv.dlevmar_chkjac.argtypes = [C.c_void_p, C.c_void_p, _dp, C.c_int, C.c_int, C.c_void_p, _dp]
            lv.dlevmar_L2nrmxmy.argtypes = [_dp, _dp, _dp, C.c_int]
            lv.dlevmar_L2nrmxmy.restype = C.c_double
            lv.dlevmar_trans_mat_mat_mult.argtypes = [_dp, _dp, C.c_int, C.c_int]

        # timing build (-O3 -march=native) of the same reference levmar + restated model: bench.py's CPU arm only
        self.ref_fast = None
        fpath = os.path.join(ORC_DIR, "_ref", "liborc_refshim_o3.so")
        if os.path.exists(fpath):
            try:
                self.ref_fast = C.CDLL(fpath, mode=os.RTLD_NOW | os.RTLD_LOCAL | getattr(os, "RTLD_DEEPBIND", 0))
                self.ref_fast.ref_fit_batch.argtypes = _FIT_ARGS
                self.ref_fast.ref_fit_batch.restype = C.c_int
            except OSError:  # built for another CPU (-march=native travels with the snapshot): fall back to the -O2 build
                self.ref_fast = None

    # ---- batched fits ---------------------------------------------------------------------------
    def fit_batch(self, model, variant, p0, meas, off, counts, opts, itmax=1000, t=None, math_mode=MATH_DET,
                  nthreads=1, use_ref=False, fast=False):
        """Runs B fits; returns dict(p, info, ret, extra).  use_ref=True routes the LM core through the
        reference's own dlevmar_der/dlevmar_dif (oracle/_ref)."""
        B = len(off) - 1
        m = 19 if model == STEREO19 else 8
        p = np.ascontiguousarray(p0, dtype=np.float64).reshape(B, m).copy()
        meas = np.ascontiguousarray(meas, dtype=np.float64)
        off = np.ascontiguousarray(off, dtype=np.int64)
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        opts_a = None if opts is None else np.ascontiguousarray(opts, dtype=np.float64)
        t_a = None if t is None else np.ascontiguousarray(t, dtype=np.float64)
        info = np.zeros((B, 10))
        ret = np.zeros(B, dtype=np.int32)
        extra = np.zeros((B, 2), dtype=np.int32)
        fn = self.ref.ref_fit_batch if use_ref else self.lib.orc_fit_batch
        if use_ref and fast and self.ref_fast is not None:
            fn = self.ref_fast.ref_fit_batch  # timing build: not bit-identical to the checker
        rc = fn(model, variant, math_mode, B, _ptr(p, _dp), _ptr(meas, _dp), _ptr(t_a, _dp), _ptr(off, _lp),
                _ptr(counts, _ip), itmax, _ptr(opts_a, _dp), _ptr(info, _dp), _ptr(ret, _ip), _ptr(extra, _ip),
                nthreads)
        if rc != 0:
            raise ValueError("oracle rejected the arguments")
        return {"p": p, "info": info, "ret": ret, "extra": extra}


    def cleanup_edges(self, pts, seg_off, tracked, reset=False):
        """Per-frame restatement of cleanup_and_group_edges; returns dict(meas, off, counts, valid) packed like the
        product's batched call."""
        pts = np.ascontiguousarray(pts, dtype=np.int32).reshape(-1, 2)
        seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
        F = (len(seg_off) - 1) // 4
        tracked = np.ascontiguousarray(tracked, dtype=np.float64).reshape(F, 2)
        counts = np.zeros((F, 4), dtype=np.int32)
        valid = np.zeros(F, dtype=np.int32)
        pieces = []
        for f in range(F):
            so = np.ascontiguousarray(seg_off[4 * f:4 * f + 5])
            rng = np.zeros(8, dtype=np.int32)
            tr = np.ascontiguousarray(tracked[f])
            ok = self.lib.orc_cleanup_edges(_ptr(pts, _ip), _ptr(so, _lp), 1 if reset else 0, _ptr(tr, _dp), _ptr(rng, _ip))
            valid[f] = ok
            if ok:
                for s in range(4):
                    c = min(int(rng[2 * s + 1] - rng[2 * s] + 1), 10000)
                    counts[f, s] = c
                    a = int(so[s]) + int(rng[2 * s])
                    pieces.append(pts[a:a + c].astype(np.float64).reshape(-1))
        meas = np.concatenate(pieces) if pieces else np.zeros(0)
        off = np.concatenate([[0], np.cumsum(2 * counts.sum(axis=1))]).astype(np.int64)
        return {"meas": meas, "off": off, "counts": counts, "valid": valid}

    def closest(self, curves, pts, is_control_points=True):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)
        Q = pts.shape[0]
        curves = np.ascontiguousarray(curves, dtype=np.float64).reshape(-1, 8)
        nearest, out, nroots = np.zeros((Q, 2)), np.zeros((Q, 2)), np.zeros(Q, dtype=np.int32)
        for q in range(Q):
            c = np.ascontiguousarray(curves[q if curves.shape[0] > 1 else 0])
            p = np.zeros(8)
            if is_control_points:
                self.lib.orc_control2coeffs(_ptr(c, _dp), _ptr(p, _dp))
            else:
                p = c
            pt = np.ascontiguousarray(pts[q])
            nroots[q] = self.lib.orc_closest_bezier_pt(_ptr(p, _dp), _ptr(pt, _dp), _ptr(nearest[q], _dp), _ptr(out[q], _dp))
        return nearest, out[:, 0].copy(), out[:, 1].copy(), nroots

    # ---- EKF ------------------------------------------------------------------------------------
    def ekf_predict(self, x, P):
        x = np.array(x, dtype=np.float64)
        P = np.array(P, dtype=np.float64, order="C")
        self.lib.orc_ekf_predict(_ptr(x, _dp), _ptr(P, _dp), x.size)
        return x, P

    def ekf_update_curves_points(self, x, P, z, A, curve_col, point_meas=None, point_col=None, want_hsk=False):
        x = np.array(x, dtype=np.float64)
        P = np.array(P, dtype=np.float64, order="C")
        N = x.size
        z = np.ascontiguousarray(z, dtype=np.float64)
        A = np.ascontiguousarray(A, dtype=np.float64).reshape(-1, 16)
        cc = np.ascontiguousarray(curve_col, dtype=np.int32)
        pm = np.zeros(0) if point_meas is None else np.ascontiguousarray(point_meas, dtype=np.float64)
        pc = np.zeros(0, dtype=np.int32) if point_col is None else np.ascontiguousarray(point_col, dtype=np.int32)
        n, n_pts = cc.size, pc.size
        M = 8 * n + 3 + 3 * n_pts
        H = np.zeros((M, N)) if want_hsk else None
        S = np.zeros((M, M)) if want_hsk else None
        K = np.zeros((N, M)) if want_hsk else None
        rc = self.lib.orc_ekf_update_curves_points(_ptr(x, _dp), _ptr(P, _dp), N, _ptr(z, _dp), n, _ptr(A, _dp),
                                                   _ptr(cc, _ip), _ptr(pm, _dp), _ptr(pc, _ip), n_pts,
                                                   _ptr(H, _dp), _ptr(S, _dp), _ptr(K, _dp))
        assert rc == 0
        return (x, P, H, S, K) if want_hsk else (x, P)

    def ekf_update_points(self, x, P, point_meas, point_col):
        x = np.array(x, dtype=np.float64)
        P = np.array(P, dtype=np.float64, order="C")
        pm = np.ascontiguousarray(point_meas, dtype=np.float64)
        pc = np.ascontiguousarray(point_col, dtype=np.int32)
        rc = self.lib.orc_ekf_update_points(_ptr(x, _dp), _ptr(P, _dp), x.size, _ptr(pm, _dp), _ptr(pc, _ip), pc.size)
        assert rc == 0
        return x, P

    def gate(self, x, P, point_col, z, gate, nthreads=1):
        x = np.ascontiguousarray(x, dtype=np.float64)
        P = np.ascontiguousarray(P, dtype=np.float64)
        pc = np.ascontiguousarray(point_col, dtype=np.int32)
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 3)
        idx = np.zeros(z.shape[0], dtype=np.int32)
        d2 = np.zeros(z.shape[0])
        rc = self.lib.orc_gate_mahalanobis(_ptr(x, _dp), _ptr(P, _dp), x.size, _ptr(pc, _ip), pc.size, _ptr(z, _dp),
                                           z.shape[0], gate, _ptr(idx, _ip), _ptr(d2, _dp), nthreads)
        assert rc == 0
        return idx, d2

    def gate_pieces(self, x, pose_rows, pose_cols, diag3, point_col, z, gate, nthreads=1, want_pack=False):
        """The gate from the pieces of P it reads: pose_rows [6, N], pose_cols [N, 6], diag3 [L, 3, 3]."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        pr = np.ascontiguousarray(pose_rows, dtype=np.float64)
        pcs = np.ascontiguousarray(pose_cols, dtype=np.float64)
        d3 = np.ascontiguousarray(diag3, dtype=np.float64)
        pc = np.ascontiguousarray(point_col, dtype=np.int32)
        assert pr.shape == (6, x.size) and pcs.shape == (x.size, 6) and d3.shape == (pc.size, 3, 3)
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 3)
        idx = np.zeros(z.shape[0], dtype=np.int32)
        d2 = np.zeros(z.shape[0])
        rc = self.lib.orc_gate_mahalanobis_pieces(_ptr(x, _dp), _ptr(pr, _dp), _ptr(pcs, _dp), _ptr(d3, _dp), x.size,
                                                  _ptr(pc, _ip), pc.size, _ptr(z, _dp), z.shape[0], gate, _ptr(idx, _ip),
                                                  _ptr(d2, _dp), nthreads)
        assert rc == 0
        if want_pack:
            pack = np.zeros((pc.size, 9))
            self.lib.orc_gate_pack_pieces(_ptr(x, _dp), _ptr(pr, _dp), _ptr(pcs, _dp), _ptr(d3, _dp), x.size,
                                          _ptr(pc, _ip), pc.size, _ptr(pack, _dp))
            return idx, d2, pack
        return idx, d2

    def gate_pack(self, x, P, point_col):
        x = np.ascontiguousarray(x, dtype=np.float64)
        P = np.ascontiguousarray(P, dtype=np.float64)
        pc = np.ascontiguousarray(point_col, dtype=np.int32)
        pack = np.zeros((pc.size, 9))
        self.lib.orc_gate_pack(_ptr(x, _dp), _ptr(P, _dp), x.size, _ptr(pc, _ip), pc.size, _ptr(pack, _dp))
        return pack


_ORACLE = None


def oracle() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE


# ---- the reference's own C++ (kalmanClass.cpp, curveFittingClass.cpp) built unchanged into oracle/_ref/libcps_ref.so
class RefCps:
    """ctypes view of oracle/ref_harness.cpp.  Present only where `make -C oracle ref` ran (needs /root/reference at
    BUILD time; the built library travels with the snapshot)."""

    def __init__(self):
        oracle()  # liborc.so and liblevmar_ref.so must be loaded (RTLD_GLOBAL) first
        path = os.path.join(ORC_DIR, "_ref", "libcps_ref.so")
        self.lib = C.CDLL(path)
        L = self.lib
        vp = C.c_void_p
        L.refh_kf_create.restype = vp
        L.refh_kf_destroy.argtypes = [vp]
        L.refh_kf_set.argtypes = [vp, _dp, _dp, C.c_int, _ip, C.c_int, _ip, C.c_int]
        L.refh_kf_dim.argtypes = [vp]
        L.refh_kf_num_curves.argtypes = [vp]
        L.refh_kf_num_points.argtypes = [vp]
        L.refh_kf_get.argtypes = [vp, _dp, _dp, _ip, _ip]
        L.refh_kf_predict.argtypes = [vp]
        L.refh_kf_add_first_states.argtypes = [vp, _dp]
        L.refh_kf_add_new_curve.argtypes = [vp, _dp, _dp]
        L.refh_kf_add_new_points.argtypes = [vp, _dp, C.c_int]
        L.refh_kf_update_curves_points.argtypes = [vp, _dp, C.c_int, _dp, _ip, _dp, _ip, C.c_int]
        L.refh_kf_update_points.argtypes = [vp, _dp, _ip, C.c_int]
        L.refh_kf_split_matrices.argtypes = [vp, C.c_double, _dp, _dp]
        L.refh_kf_predicted_measurement.argtypes = [vp, _dp, C.c_int, _dp]
        L.refh_kf_rbe_derivs.argtypes = [vp, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp]
        L.refh_generate_rbe.argtypes = [C.c_double, C.c_double, C.c_double, _dp]
        L.refh_modelfunc.argtypes = [_dp, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.refh_cp_earth2image.argtypes = [_dp, _dp, _dp, _dp, _dp]
        L.refh_earth2body.argtypes = [_dp, _dp, _dp, _dp]
        L.refh_bezier_eval.argtypes = [_dp, C.c_double, _dp, _dp]
        L.refh_bezier_eval.restype = C.c_double
        L.refh_closest_bezier_pt.argtypes = [_dp, _dp, _dp, _dp]
        L.refh_control2coeffs.argtypes = [_dp, _dp]
        L.refh_fit_curve.argtypes = [_dp, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int]
        L.refh_fit_curve.restype = C.c_double
        L.refh_last_fit.argtypes = [_dp, _dp, _dp, _ip, _dp, _ip, _dp, _dp, _ip]
        L.refh_cleanup_and_group_edges.argtypes = [_ip, _lp, C.c_int, _ip, _dp, _ip, _dp, C.c_longlong]

    # -- KalmanFilter -----------------------------------------------------------------------------------------
    class KF:
        def __init__(self, ref):
            self.r = ref
            self.h = C.c_void_p(ref.lib.refh_kf_create())

        def close(self):
            if self.h:
                self.r.lib.refh_kf_destroy(self.h)
                self.h = None

        def set(self, x, P, curve_inds, point_inds):
            x = np.ascontiguousarray(x, dtype=np.float64)
            P = np.ascontiguousarray(P, dtype=np.float64)
            ci = np.ascontiguousarray(curve_inds, dtype=np.int32)
            pi = np.ascontiguousarray(point_inds, dtype=np.int32)
            assert x.size == 12 + 8 * ci.size + 3 * pi.size, "the reference derives N from num_curves / num_points"
            self.r.lib.refh_kf_set(self.h, _ptr(x, _dp), _ptr(P, _dp), x.size, _ptr(ci, _ip), ci.size, _ptr(pi, _ip),
                                   pi.size)

        def get(self):
            L = self.r.lib
            N = L.refh_kf_dim(self.h)
            nc, npt = L.refh_kf_num_curves(self.h), L.refh_kf_num_points(self.h)
            x, P = np.zeros(N), np.zeros((N, N))
            ci, pi = np.zeros(max(nc, 1), dtype=np.int32), np.zeros(max(npt, 1), dtype=np.int32)
            L.refh_kf_get(self.h, _ptr(x, _dp), _ptr(P, _dp), _ptr(ci, _ip), _ptr(pi, _ip))
            return x, P, ci[:nc], pi[:npt]

        def predict(self):
            self.r.lib.refh_kf_predict(self.h)

        def add_first_states(self, z19):
            z = np.ascontiguousarray(z19, dtype=np.float64)
            self.r.lib.refh_kf_add_first_states(self.h, _ptr(z, _dp))

        def add_new_curve(self, z8, A):
            z = np.ascontiguousarray(z8, dtype=np.float64)
            A = np.ascontiguousarray(A, dtype=np.float64)
            self.r.lib.refh_kf_add_new_curve(self.h, _ptr(z, _dp), _ptr(A, _dp))

        def add_new_points(self, xyz):
            m = np.ascontiguousarray(xyz, dtype=np.float64)
            self.r.lib.refh_kf_add_new_points(self.h, _ptr(m, _dp), m.size // 3)

        def update_curves_points(self, z, A, curve_num, point_meas=None, point_nums=None):
            z = np.ascontiguousarray(z, dtype=np.float64)
            A = np.ascontiguousarray(A, dtype=np.float64).reshape(-1, 16)
            cn = np.ascontiguousarray(curve_num, dtype=np.int32)
            pm = np.zeros(1) if point_meas is None or len(point_meas) == 0 else np.ascontiguousarray(point_meas, dtype=np.float64)
            pn = np.zeros(1, dtype=np.int32) if point_nums is None or len(point_nums) == 0 else np.ascontiguousarray(point_nums, dtype=np.int32)
            n_pts = 0 if point_nums is None else len(point_nums)
            self.r.lib.refh_kf_update_curves_points(self.h, _ptr(z, _dp), cn.size, _ptr(A, _dp), _ptr(cn, _ip),
                                                    _ptr(pm, _dp), _ptr(pn, _ip), n_pts)

        def update_points(self, point_meas, point_nums):
            pm = np.ascontiguousarray(point_meas, dtype=np.float64)
            pn = np.ascontiguousarray(point_nums, dtype=np.int32)
            self.r.lib.refh_kf_update_points(self.h, _ptr(pm, _dp), _ptr(pn, _ip), pn.size)

        def split_matrices(self, t):
            A1, A2 = np.zeros((4, 4)), np.zeros((4, 4))
            self.r.lib.refh_kf_split_matrices(self.h, float(t), _ptr(A1, _dp), _ptr(A2, _dp))
            return A1, A2

        def predicted_measurement(self, A, num_curve):
            A = np.ascontiguousarray(A, dtype=np.float64)
            zh = np.zeros(8)
            self.r.lib.refh_kf_predicted_measurement(self.h, _ptr(A, _dp), int(num_curve), _ptr(zh, _dp))
            return zh

    def kf(self):
        return RefCps.KF(self)

    # -- curve model / driver ---------------------------------------------------------------------------------
    def modelfunc(self, p19, t, counts):
        p = np.ascontiguousarray(p19, dtype=np.float64)
        t = np.ascontiguousarray(t, dtype=np.float64)
        n = 2 * int(np.sum(counts))
        assert t.size == n // 2
        hx = np.zeros(n)
        l1, l2, r1, r2 = (int(c) for c in counts)  # measurement order [L c1 | L c2 | R c1 | R c2]
        self.lib.refh_modelfunc(_ptr(p, _dp), _ptr(hx, _dp), n, _ptr(t, _dp), l1, r1, l2, r2)
        return hx

    def generate_rbe(self, phi, theta, psi):
        R = np.zeros((3, 3))
        self.lib.refh_generate_rbe(float(phi), float(theta), float(psi), _ptr(R, _dp))
        return R

    def rbe_derivs(self, phi, theta, psi):
        kf = self.kf()
        a, b, c = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 3))
        self.lib.refh_kf_rbe_derivs(kf.h, float(phi), float(theta), float(psi), _ptr(a, _dp), _ptr(b, _dp), _ptr(c, _dp))
        kf.close()
        return a, b, c

    def cp_earth2image(self, c_earth8, euler, translation):
        c = np.ascontiguousarray(c_earth8, dtype=np.float64)
        e = np.ascontiguousarray(euler, dtype=np.float64)
        tr = np.ascontiguousarray(translation, dtype=np.float64)
        l, r = np.zeros(8), np.zeros(8)
        self.lib.refh_cp_earth2image(_ptr(c, _dp), _ptr(l, _dp), _ptr(r, _dp), _ptr(e, _dp), _ptr(tr, _dp))
        return l, r

    def bezier_eval(self, cp8, t):
        cp = np.ascontiguousarray(cp8, dtype=np.float64)
        pt, near = np.zeros(2), np.zeros(2)
        self.lib.refh_bezier_eval(_ptr(cp, _dp), float(t), _ptr(pt, _dp), _ptr(near, _dp))
        return near

    def fit_curve(self, meas, counts):
        """The reference's CurveFittingClass::fit_curve on one frame; meas packed [L c1 | L c2 | R c1 | R c2]."""
        meas = np.ascontiguousarray(meas, dtype=np.float64)
        l1, l2, r1, r2 = (int(c) for c in counts)
        o = np.cumsum([0, 2 * l1, 2 * l2, 2 * r1, 2 * r2])
        seg = [np.ascontiguousarray(meas[o[i]:o[i + 1]]) for i in range(4)]
        p = np.zeros(19)
        err = self.lib.refh_fit_curve(_ptr(p, _dp), _ptr(seg[0], _dp), l1, _ptr(seg[1], _dp), l2, _ptr(seg[2], _dp), r1,
                                      _ptr(seg[3], _dp), r2)
        return p, err

    def last_fit(self):
        """What the last fit_curve handed to levmar (initial guess, packed measurements, t, counts, opts, itmax) and
        what levmar returned before the post-normalisation (p_lm, info, ret)."""
        n = self.lib.refh_last_fit(None, None, None, None, None, None, None, None, None)
        p0, x, t, p_lm, info, opts = np.zeros(19), np.zeros(n), np.zeros(n // 2), np.zeros(19), np.zeros(10), np.zeros(5)
        counts, itmax, ret = np.zeros(4, dtype=np.int32), C.c_int(0), C.c_int(0)
        self.lib.refh_last_fit(_ptr(p0, _dp), _ptr(x, _dp), _ptr(t, _dp), _ptr(counts, _ip), _ptr(opts, _dp),
                               C.byref(itmax), _ptr(p_lm, _dp), _ptr(info, _dp), C.byref(ret))
        return dict(p0=p0, x=x, t=t, counts=counts, opts=opts, itmax=itmax.value, p_lm=p_lm, info=info, ret=ret.value)

    def closest(self, p8, pt):
        p = np.ascontiguousarray(p8, dtype=np.float64)
        q = np.ascontiguousarray(pt, dtype=np.float64)
        near, out = np.zeros(2), np.zeros(2)
        self.lib.refh_closest_bezier_pt(_ptr(p, _dp), _ptr(q, _dp), _ptr(near, _dp), _ptr(out, _dp))
        return near, out[0], out[1]

    def control2coeffs(self, cp8):
        c = np.ascontiguousarray(cp8, dtype=np.float64)
        p = np.zeros(8)
        self.lib.refh_control2coeffs(_ptr(c, _dp), _ptr(p, _dp))
        return p

    def cleanup_edges(self, pts, seg_off, tracked, reset=False):
        pts = np.ascontiguousarray(pts, dtype=np.int32).reshape(-1, 2)
        so = np.ascontiguousarray(seg_off, dtype=np.int64)
        tr = np.ascontiguousarray(tracked, dtype=np.float64).copy()
        cut = np.zeros(2, dtype=np.int32)
        counts = np.zeros(4, dtype=np.int32)
        cap = 2 * int(so[4] - so[0]) + 8
        meas = np.zeros(cap)
        ok = self.lib.refh_cleanup_and_group_edges(_ptr(pts, _ip), _ptr(so, _lp), 1 if reset else 0, _ptr(cut, _ip),
                                                   _ptr(tr, _dp), _ptr(counts, _ip), _ptr(meas, _dp), cap)
        return ok, counts, meas[:2 * int(counts.sum())], cut, tr


_REFCPS = None


def refcps():
    """The reference build, or None where oracle/_ref/libcps_ref.so does not exist."""
    global _REFCPS
    if _REFCPS is None:
        if not os.path.exists(os.path.join(ORC_DIR, "_ref", "libcps_ref.so")) or oracle().levmar is None:
            return None
        _REFCPS = RefCps()
    return _REFCPS
