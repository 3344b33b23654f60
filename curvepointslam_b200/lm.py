"""Host-side mirror of the reference's curve-fitting operator surface over libcps.so.

Reference interface mirrored here (names, argument meaning, error behaviour):
  * levmar's ``dlevmar_der`` / ``dlevmar_dif`` (levmar-2.5/levmar.h:98-107) -> batched calls with the same
    ``p`` in/out, ``opts``, ``info[10]`` and return-value conventions;
  * ``CurveFittingClass::fit_curve`` (curveFittingClass.cpp:408-724; declaration curveFittingClass.h:82):
    the reference's live options ``{1e-10, 1e-10, 1e-10, 1e-20, 1e-5}``, ``itmax=1000``, ``dlevmar_dif``
    (:661-681), ``fitting_error = info[1]`` (:683) and the post-fit sign/angle normalisation (:684-711).
All arithmetic of the fits happens in the CUDA kernels; this file only moves pointers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

STEREO19, IMAGE8 = 19, 8
DER, DIF = 0, 1
REFERENCE_OPTS = (1e-10, 1e-10, 1e-10, 1e-20, 1e-6 * 1e1)  # curveFittingClass.cpp:661-662
REFERENCE_ITMAX = 1000  # curveFittingClass.cpp:681
PATH_AUTO, PATH_MONOLITHIC, PATH_STAGED = 0, 1, 2
PI = 3.1415926535  # common.h:114


def _vp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class LmContext:
    """One CUDA stream + workspace (cps_lm_ctx).  Not thread-safe; make one per thread."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _lib.check(_lib.lib().cps_lm_create(C.byref(self._h), device), "cps_lm_create")
        self.device = device

    def close(self):
        if self._h:
            _lib.lib().cps_lm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_path(self, path: int):
        """Execution shape of dlevmar_der on Stereo19 batches (cps_lm_set_path): PATH_AUTO, PATH_MONOLITHIC (a warp
        per fit, one kernel) or PATH_STAGED (rounds of phase kernels, a thread per fit).  Same results bit for bit."""
        _lib.check(_lib.lib().cps_lm_set_path(self._h, int(path)), "cps_lm_set_path")

    def set_profiling(self, on: bool):
        _lib.check(_lib.lib().cps_lm_set_profiling(self._h, 1 if on else 0), "cps_lm_set_profiling")

    def get_profile(self, reset: bool = True):
        """Accumulated CUDA-event time and launch count of the staged shape's kernels: {kernel: (ms, launches)}."""
        ms = (C.c_double * 4)()
        n = (C.c_longlong * 4)()
        _lib.check(_lib.lib().cps_lm_get_profile(self._h, ms, n, 1 if reset else 0), "cps_lm_get_profile")
        return {k: (ms[i], n[i]) for i, k in enumerate(("solve", "eval", "jac", "tail"))}

    def get_profile_work(self, reset: bool = True):
        """(Jacobian evaluations staged_jac_kernel performed, fits the tail kernel took over) while profiling was on."""
        je, tf = C.c_longlong(0), C.c_longlong(0)
        _lib.check(_lib.lib().cps_lm_get_profile_work(self._h, C.byref(je), C.byref(tf), 1 if reset else 0),
                   "cps_lm_get_profile_work")
        return int(je.value), int(tf.value)

    def set_async(self, on: bool):
        """Asynchronous device entry: fit_batched_dev queues the call for a worker thread and makes `stream` wait for it."""
        _lib.check(_lib.lib().cps_lm_set_async(self._h, 1 if on else 0), "cps_lm_set_async")

    def set_arith(self, fma: bool):
        """Fused multiply-adds in the staged dlevmar_der kernels (not bit-identical to the reference; default exact)."""
        _lib.check(_lib.lib().cps_lm_set_arith(self._h, 1 if fma else 0), "cps_lm_set_arith")

    def solve_handovers(self) -> int:
        """Systems of the last staged dlevmar_der call that the block-arrow solve handed to the dense one."""
        n = C.c_longlong(0)
        _lib.check(_lib.lib().cps_lm_get_solve_handovers(self._h, C.byref(n)), "cps_lm_get_solve_handovers")
        return int(n.value)

    # -- host buffers ---------------------------------------------------------------------------------
    def fit_batched(self, model: int, variant: int, p0, meas, off, counts, opts=REFERENCE_OPTS,
                    itmax: int = REFERENCE_ITMAX, t=None):
        """B fits from host arrays.  Returns dict(p [B,m], info [B,10], ret [B], extra [B,2])."""
        off = np.ascontiguousarray(off, dtype=np.int64)
        B = len(off) - 1
        m = 19 if model == STEREO19 else 8
        p = np.ascontiguousarray(p0, dtype=np.float64).reshape(B, m).copy()
        meas = np.ascontiguousarray(meas, dtype=np.float64)
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        t_a = None if t is None else np.ascontiguousarray(t, dtype=np.float64)
        opts_a = None if opts is None else np.ascontiguousarray(opts, dtype=np.float64)
        info = np.zeros((B, 10))
        ret = np.zeros(B, dtype=np.int32)
        extra = np.zeros((B, 2), dtype=np.int32)
        fn = _lib.lib().cps_dlevmar_der_batched if variant == DER else _lib.lib().cps_dlevmar_dif_batched
        rc = fn(self._h, model, B, _vp(p), _vp(meas), _vp(t_a), _vp(off), _vp(counts), itmax, _vp(opts_a),
                _vp(info), _vp(ret), _vp(extra))
        _lib.check(rc, "cps_dlevmar_*_batched")
        return {"p": p, "info": info, "ret": ret, "extra": extra}

    def fit_batched_pix(self, model: int, variant: int, p0, pix, off, counts, opts=REFERENCE_OPTS,
                        itmax: int = REFERENCE_ITMAX):
        """The same with the samples as integer pixel coordinates (int32 or int16 array, measurement order): 4 or 2
        bytes per coordinate cross the bus instead of 8."""
        off = np.ascontiguousarray(off, dtype=np.int64)
        B = len(off) - 1
        m = 19 if model == STEREO19 else 8
        p = np.ascontiguousarray(p0, dtype=np.float64).reshape(B, m).copy()
        pix = np.ascontiguousarray(pix)
        if pix.dtype not in (np.int32, np.int16):
            raise TypeError("pix must be int32 or int16")
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        opts_a = None if opts is None else np.ascontiguousarray(opts, dtype=np.float64)
        info = np.zeros((B, 10))
        ret = np.zeros(B, dtype=np.int32)
        extra = np.zeros((B, 2), dtype=np.int32)
        rc = _lib.lib().cps_dlevmar_batched_pix(self._h, model, variant, B, _vp(p), _vp(pix), 1 if pix.dtype == np.int16 else 0,
                                               _vp(off), _vp(counts), itmax, _vp(opts_a), _vp(info), _vp(ret), _vp(extra))
        _lib.check(rc, "cps_dlevmar_batched_pix")
        return {"p": p, "info": info, "ret": ret, "extra": extra}

    # -- device buffers (raw pointers, e.g. torch tensors' data_ptr()) ---------------------------------
    def fit_batched_dev(self, model: int, variant: int, B: int, d_p: int, d_meas: int, d_t: int | None,
                        d_off: int, d_counts: int, n_max: int, d_info: int, d_ret: int, d_extra: int | None,
                        opts=REFERENCE_OPTS, itmax: int = REFERENCE_ITMAX, stream: int | None = None):
        opts_a = None if opts is None else np.ascontiguousarray(opts, dtype=np.float64)
        rc = _lib.lib().cps_lm_fit_batched_dev(self._h, model, variant, B, d_p, d_meas, d_t, d_off, d_counts,
                                               n_max, itmax, _vp(opts_a), d_info, d_ret, d_extra, stream)
        _lib.check(rc, "cps_lm_fit_batched_dev")

    def init_guess(self, meas, off, counts):
        """fit_curve's initial guess (curveFittingClass.cpp:516-651) for B Stereo19 fits -> p [B,19]."""
        off = np.ascontiguousarray(off, dtype=np.int64)
        B = len(off) - 1
        meas = np.ascontiguousarray(meas, dtype=np.float64)
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        p = np.zeros((B, 19))
        _lib.check(_lib.lib().cps_fit_curve_init_batched(self._h, B, _vp(meas), _vp(off), _vp(counts), _vp(p)),
                   "cps_fit_curve_init_batched")
        return p

    def cleanup_and_group_edges(self, pts, seg_off, tracked_endpt_y, reset_data_assoc=False):
        """CurveFittingClass::cleanup_and_group_edges (curveFittingClass.cpp:1207-1418) for F frames.
        pts [P,2] int32 pixel (x,y); seg_off [4F+1] point offsets ([L0, L1, R0, R1] per frame); tracked_endpt_y [F,2].
        Returns dict(meas, off, counts, valid): the inputs of fit_batched / init_guess."""
        pts = np.ascontiguousarray(pts, dtype=np.int32).reshape(-1, 2)
        seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
        F = (len(seg_off) - 1) // 4
        tr = np.ascontiguousarray(tracked_endpt_y, dtype=np.float64).reshape(F, 2)
        meas = np.zeros(2 * max(len(pts), 1))
        off = np.zeros(F + 1, dtype=np.int64)
        counts = np.zeros((F, 4), dtype=np.int32)
        valid = np.zeros(F, dtype=np.int32)
        _lib.check(_lib.lib().cps_cleanup_and_group_edges_batched(self._h, F, _vp(pts), _vp(seg_off),
                                                                  1 if reset_data_assoc else 0, _vp(tr), _vp(meas),
                                                                  _vp(off), _vp(counts), _vp(valid)),
                   "cps_cleanup_and_group_edges_batched")
        return {"meas": meas[: off[F]], "off": off, "counts": counts, "valid": valid}

    def fit_frames(self, variant: int, pts, seg_off, tracked_endpt_y, reset_data_assoc=False, opts=REFERENCE_OPTS,
                   itmax: int = REFERENCE_ITMAX):
        """cps_fit_frames_batched: cleanup_and_group_edges -> fit_curve for F frames in one device-resident pass.
        pts: [P,2] int32 or int16 pixel pairs.  Returns dict(p, info, ret, valid, fitting_error)."""
        pts = np.ascontiguousarray(pts)
        if pts.dtype not in (np.int16, np.int32):
            pts = pts.astype(np.int32)
        pts = pts.reshape(-1, 2)
        so = np.ascontiguousarray(seg_off, dtype=np.int64)
        F = (so.size - 1) // 4
        tr = np.ascontiguousarray(tracked_endpt_y, dtype=np.float64).reshape(F, 2)
        opts_a = None if opts is None else np.ascontiguousarray(opts, dtype=np.float64)
        p = np.zeros((F, 19))
        info = np.zeros((F, 10))
        ret = np.zeros(F, dtype=np.int32)
        valid = np.zeros(F, dtype=np.int32)
        err = np.zeros(F)
        rc = _lib.lib().cps_fit_frames_batched(self._h, variant, F, _vp(pts), 1 if pts.dtype == np.int16 else 0, _vp(so),
                                               1 if reset_data_assoc else 0, _vp(tr), itmax, _vp(opts_a), _vp(p), _vp(info),
                                               _vp(ret), _vp(valid), _vp(err))
        _lib.check(rc, "cps_fit_frames_batched")
        return {"p": p, "info": info, "ret": ret, "valid": valid, "fitting_error": err}

    def sync(self):
        _lib.check(_lib.lib().cps_lm_sync(self._h), "cps_lm_sync")

    def launch_stats(self):
        g, b, s, n = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
        _lib.lib().cps_lm_launch_stats(self._h, C.byref(g), C.byref(b), C.byref(s), C.byref(n))
        return {"grid": g.value, "block": b.value, "smem_bytes": s.value, "launches": n.value}


def closest_bezier_pt(curves, pts, is_control_points=True, device: int = 0):
    """closest_bezier_pt(p, pt, nearest_pt, out) of the reference (curveFittingClass.cpp:936-1043) for Q queries.
    curves: [Q,8] (or [8] shared by all queries); pts [Q,2].  Returns (nearest [Q,2], dist [Q], t [Q], nroots [Q])."""
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)
    Q = pts.shape[0]
    curves = np.ascontiguousarray(curves, dtype=np.float64)
    shared = 1 if curves.size == 8 else 0
    nearest, out = np.zeros((Q, 2)), np.zeros((Q, 2))
    nroots = np.zeros(Q, dtype=np.int32)
    _lib.check(_lib.lib().cps_closest_bezier_pt_batched(device, Q, _vp(curves), 1 if is_control_points else 0, shared,
                                                        _vp(pts), _vp(nearest), _vp(out), _vp(nroots)),
               "cps_closest_bezier_pt_batched")
    return nearest, out[:, 0].copy(), out[:, 1].copy(), nroots


def postfit_normalise(p: np.ndarray) -> np.ndarray:
    """fit_curve's post-fit sign/angle normalisation (curveFittingClass.cpp:684-711), vectorised over fits;
    including the reference's ``params[18] += PI`` in the second branch."""
    p = np.array(p, dtype=np.float64, copy=True).reshape(-1, 19)
    flip = p[:, 18] > 0.0
    p[flip, 1:16:2] *= -1.0
    p[flip, 16] *= -1.0
    p[flip, 17] += PI
    p[flip, 18] *= -1.0
    b2 = p[:, 16] < -PI / 2.0
    p[b2, 1:16:2] *= -1.0
    p[b2, 17] = -PI - p[b2, 17]
    p[b2, 18] += PI
    for col in (16, 17):
        for _ in range(64):
            hi = p[:, col] > PI
            lo = p[:, col] < -PI
            if not (hi.any() or lo.any()):
                break
            p[hi, col] -= 2 * PI
            p[lo, col] += 2 * PI
    return p


class CurveFittingClass:
    """Mirror of the reference class for the fit path (curveFittingClass.h:77-121).

    ``fit_curve(L, R)`` takes the two left-image and two right-image contour segments as [k,2] arrays of
    (x, y) pixels (the reference's 2k x 1 interleaved ``CvMat``s, curveFittingClass.cpp:435-489) and the
    initial parameters; ``fit_curves_batched`` is the same for many frames at once."""

    def __init__(self, ctx: LmContext | None = None, variant: int = DIF):
        self.ctx = ctx or LmContext()
        self.variant = variant
        self.fitting_error = 0.0
        self.info = None

    def fit_curves_batched(self, params0, frames):
        """frames: list of (L0, L1, R0, R1) arrays [k,2]; params0 [B,19] initial parameters."""
        B = len(frames)
        pieces, counts = [], np.zeros((B, 4), dtype=np.int32)
        for b, (l0, l1, r0, r1) in enumerate(frames):
            for s, seg in enumerate((l0, l1, r0, r1)):  # [L c1 | L c2 | R c1 | R c2]
                a = np.asarray(seg, dtype=np.float64).reshape(-1, 2)
                counts[b, s] = a.shape[0]
                pieces.append(a.reshape(-1))
        meas = np.concatenate(pieces) if pieces else np.zeros(0)
        off = np.concatenate([[0], np.cumsum(2 * counts.sum(axis=1))]).astype(np.int64)
        if params0 is None:  # the reference ignores the incoming params and builds its own guess (:516-651)
            params0 = self.ctx.init_guess(meas, off, counts)
        r = self.ctx.fit_batched(STEREO19, self.variant, params0, meas, off, counts)
        self.info = r["info"]
        return postfit_normalise(r["p"]), r["info"][:, 1].copy(), r

    def fit_curve(self, params0, L, R):
        p0 = None if params0 is None else np.asarray(params0, dtype=np.float64).reshape(1, 19)
        p, err, _ = self.fit_curves_batched(p0, [(L[0], L[1], R[0], R[1])])
        self.fitting_error = float(err[0])
        return p[0]
