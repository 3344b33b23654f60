"""Host-side mirror of the reference's ``KalmanFilter`` (kalmanClass.h:40-163) for the hot path, over the
C ABI of include/cps_ekf.h.  Method names and argument meaning follow the reference:

    PredictKF()                                                    kalmanClass.cpp:161-272
    UpdateNCurvesAndPoints(z, n, A, curve_num, point_meas, point_nums, n_pts)   :669-945
    UpdatePoints(point_meas, point_nums, n_pts)                    :948-1126
    GetSplitMatrices(t) -> (A1, A2)                                :1144-1159
    getState() -> x                                                :1138-1141
    members x, P (device resident: read back through getState()/getCov()), num_curves, num_points,
    curve_inds, point_inds                                         kalmanClass.h:152-162

plus ``gate(z, gate)``, the Mahalanobis gate this project defines for the point landmarks (the reference
has none).  All arithmetic runs in the CUDA kernels; nothing here computes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

ROBOT_STATE_SIZE = 12
CHI2_3DOF_99 = 11.345  # default gate: chi-square, 3 dof, 0.99


def _bind():
    L = _lib.lib()
    if getattr(L, "_ekf_bound", False):
        return L
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    L.cps_ekf_create.argtypes = [C.POINTER(vp), ci, ci]
    L.cps_ekf_destroy.argtypes = [vp]
    L.cps_ekf_destroy.restype = None
    L.cps_ekf_sync.argtypes = [vp]
    L.cps_ekf_set_state.argtypes = [vp, vp, vp, ci, vp, ci, vp, ci]
    L.cps_ekf_set_cov_lowrank.argtypes = [vp, vp, ci, cd, cd]
    L.cps_ekf_dims.argtypes = [vp, vp, vp, vp]
    L.cps_ekf_get_state.argtypes = [vp, vp]
    L.cps_ekf_get_cov.argtypes = [vp, vp]
    L.cps_ekf_get_cov_block.argtypes = [vp, ci, ci, ci, ci, vp]
    L.cps_ekf_predict.argtypes = [vp]
    L.cps_ekf_add_first_states.argtypes = [vp, vp]
    L.cps_ekf_add_new_curve.argtypes = [vp, vp, vp]
    L.cps_ekf_add_new_points.argtypes = [vp, vp, ci]
    L.cps_ekf_get_index_tables.argtypes = [vp, vp, vp]
    L.cps_ekf_update_curves_points.argtypes = [vp, vp, ci, vp, vp, vp, vp, ci]
    L.cps_ekf_update_points.argtypes = [vp, vp, vp, ci]
    L.cps_ekf_last_timing.argtypes = [vp, vp, vp, vp]
    L.cps_ekf_split_matrices.argtypes = [cd, vp, vp]
    L.cps_ekf_split_matrices.restype = None
    L.cps_gate_mahalanobis.argtypes = [vp, vp, ci, cd, vp, vp]
    L.cps_gate_mahalanobis_dev.argtypes = [vp, vp, ci, cd, vp, vp]
    L.cps_gate_last_timing.argtypes = [vp, vp, vp]
    L._ekf_bound = True
    return L


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class KalmanFilter:
    def __init__(self, capacity: int, device: int = 0):
        self._L = _bind()
        self._h = C.c_void_p()
        _lib.check(self._L.cps_ekf_create(C.byref(self._h), device, int(capacity)), "cps_ekf_create")
        self.curve_inds: list[int] = []
        self.point_inds: list[int] = []

    def close(self):
        if getattr(self, "_h", None):
            self._L.cps_ekf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state in / out ------------------------------------------------------------------------------
    def set_state(self, x, P, curve_inds=(), point_inds=()):
        x = np.ascontiguousarray(x, dtype=np.float64)
        Pm = None if P is None else np.ascontiguousarray(P, dtype=np.float64)
        ci = np.ascontiguousarray(curve_inds, dtype=np.int32)
        pi = np.ascontiguousarray(point_inds, dtype=np.int32)
        _lib.check(self._L.cps_ekf_set_state(self._h, _p(x), _p(Pm), x.size, _p(ci), ci.size, _p(pi), pi.size),
                   "cps_ekf_set_state")
        self.curve_inds, self.point_inds = list(map(int, ci)), list(map(int, pi))

    def set_cov_lowrank(self, G, scale, diag):
        G = np.ascontiguousarray(G, dtype=np.float64)
        _lib.check(self._L.cps_ekf_set_cov_lowrank(self._h, _p(G), G.shape[1], scale, diag), "cps_ekf_set_cov_lowrank")

    @property
    def num_curves(self):
        return len(self.curve_inds)

    @property
    def num_points(self):
        return len(self.point_inds)

    @property
    def N(self):
        n = C.c_int()
        self._L.cps_ekf_dims(self._h, C.byref(n), None, None)
        return n.value

    def getState(self):
        x = np.zeros(self.N)
        _lib.check(self._L.cps_ekf_get_state(self._h, _p(x)), "cps_ekf_get_state")
        return x

    def getCov(self):
        n = self.N
        P = np.zeros((n, n))
        _lib.check(self._L.cps_ekf_get_cov(self._h, _p(P)), "cps_ekf_get_cov")
        return P

    def getCovBlock(self, r0, c0, nr, nc):
        out = np.zeros((nr, nc))
        _lib.check(self._L.cps_ekf_get_cov_block(self._h, r0, c0, nr, nc, _p(out)), "cps_ekf_get_cov_block")
        return out

    x = property(getState)
    P = property(getCov)

    # ---- the reference's entry points ------------------------------------------------------------------
    def PredictKF(self):
        _lib.check(self._L.cps_ekf_predict(self._h), "cps_ekf_predict")

    def UpdateNCurvesAndPoints(self, measurement, n, A, curve_num, point_meas=None, point_nums=None, n_pts=0):
        z = np.ascontiguousarray(measurement, dtype=np.float64).reshape(-1)
        Aa = np.ascontiguousarray(A, dtype=np.float64).reshape(-1) if n else None
        cn = np.ascontiguousarray(curve_num, dtype=np.int32) if n else None
        pm = np.ascontiguousarray(point_meas, dtype=np.float64).reshape(-1) if n_pts else None
        pn = np.ascontiguousarray(point_nums, dtype=np.int32) if n_pts else None
        _lib.check(self._L.cps_ekf_update_curves_points(self._h, _p(z), n, _p(Aa), _p(cn), _p(pm), _p(pn), n_pts),
                   "cps_ekf_update_curves_points")

    def UpdatePoints(self, point_meas, point_nums, n_pts):
        pm = np.ascontiguousarray(point_meas, dtype=np.float64).reshape(-1)
        pn = np.ascontiguousarray(point_nums, dtype=np.int32)
        _lib.check(self._L.cps_ekf_update_points(self._h, _p(pm), _p(pn), n_pts), "cps_ekf_update_points")

    # ---- state augmentation (kalmanClass.cpp:275-666), in place on the device ---------------------------
    def _refresh_tables(self):
        nc, npnt = C.c_int(), C.c_int()
        self._L.cps_ekf_dims(self._h, None, C.byref(nc), C.byref(npnt))
        ci, pi = np.zeros(max(nc.value, 1), dtype=np.int32), np.zeros(max(npnt.value, 1), dtype=np.int32)
        self._L.cps_ekf_get_index_tables(self._h, _p(ci), _p(pi))
        self.curve_inds, self.point_inds = list(map(int, ci[:nc.value])), list(map(int, pi[:npnt.value]))

    def AddFirstStates(self, measurement):
        z = np.ascontiguousarray(measurement, dtype=np.float64).reshape(-1)
        _lib.check(self._L.cps_ekf_add_first_states(self._h, _p(z)), "cps_ekf_add_first_states")
        self._refresh_tables()

    def AddNewCurve(self, z, A):
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1)
        Aa = np.ascontiguousarray(A, dtype=np.float64).reshape(-1)
        _lib.check(self._L.cps_ekf_add_new_curve(self._h, _p(z), _p(Aa)), "cps_ekf_add_new_curve")
        self._refresh_tables()

    def AddNewPoints(self, measurements, n_pts):
        m = np.ascontiguousarray(measurements, dtype=np.float64).reshape(-1)
        _lib.check(self._L.cps_ekf_add_new_points(self._h, _p(m), int(n_pts)), "cps_ekf_add_new_points")
        self._refresh_tables()

    def GetSplitMatrices(self, t):
        A1, A2 = np.zeros((4, 4)), np.zeros((4, 4))
        self._L.cps_ekf_split_matrices(float(t), _p(A1), _p(A2))
        return A1, A2

    def sync(self):
        _lib.check(self._L.cps_ekf_sync(self._h), "cps_ekf_sync")

    def last_timing(self):
        t, c, n = C.c_float(), C.c_float(), C.c_longlong()
        _lib.check(self._L.cps_ekf_last_timing(self._h, C.byref(t), C.byref(c), C.byref(n)), "cps_ekf_last_timing")
        return {"total_ms": t.value, "cov_kernel_ms": c.value, "launches": n.value}

    # ---- Mahalanobis gate ----------------------------------------------------------------------------
    def gate(self, z, gate=CHI2_3DOF_99, want_d2=True):
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 3)
        idx = np.zeros(z.shape[0], dtype=np.int32)
        d2 = np.zeros(z.shape[0]) if want_d2 else None
        _lib.check(self._L.cps_gate_mahalanobis(self._h, _p(z), z.shape[0], float(gate), _p(idx), _p(d2)),
                   "cps_gate_mahalanobis")
        return (idx, d2) if want_d2 else idx

    def gate_dev(self, d_z: int, nz: int, gate: float, d_idx: int, d_d2: int | None):
        _lib.check(self._L.cps_gate_mahalanobis_dev(self._h, d_z, nz, float(gate), d_idx, d_d2),
                   "cps_gate_mahalanobis_dev")

    def gate_timing(self):
        a, b = C.c_float(), C.c_float()
        _lib.check(self._L.cps_gate_last_timing(self._h, C.byref(a), C.byref(b)), "cps_gate_last_timing")
        return {"pack_ms": a.value, "sweep_ms": b.value}


def make_synthetic_ekf(L: int, d: int = 3, seed: int = 0, rank: int = 64):
    """SURVEY.md 8(d) EKF workload: N = 12 + d*L; pose (0,0,-1; 0.01,-0.02,0.3; 1,0,0; 0,0,0); landmarks
    U(-50,50) (curves: ground z = 0 means only x/y control points); P = (1/rank) G G^T + 0.25 I with G N x rank
    standard normal.  Returns x, G, curve_inds, point_inds (P is formed on the device from G)."""
    rng = np.random.Generator(np.random.Philox(key=[seed, 77]))
    N = 12 + d * L
    x = np.zeros(N)
    x[:12] = (0.0, 0.0, -1.0, 0.01, -0.02, 0.3, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    x[12:] = rng.uniform(-50.0, 50.0, N - 12)
    G = rng.standard_normal((N, rank))
    if d == 8:
        return x, G, np.arange(L, dtype=np.int32) * 8 + 12, np.zeros(0, dtype=np.int32)
    # d == 3: point landmarks, plus 4 curves carved out of the first 32 landmark states so that the
    # reference's live measurement (n = 4 curves, M = 35) has something to update
    n_c = 4
    curve_inds = 12 + 8 * np.arange(n_c, dtype=np.int32)
    first_pt = 12 + 8 * n_c
    n_p = (N - first_pt) // 3
    point_inds = first_pt + 3 * np.arange(n_p, dtype=np.int32)
    return x, G, curve_inds, point_inds
