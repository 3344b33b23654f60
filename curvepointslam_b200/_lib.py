"""ctypes binding of libcps.so.  There is no fallback: if the library is missing or has no CUDA device
to run on, the calls raise."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CPS_LIB", os.path.join(HERE, "libcps.so"))

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
lp = C.POINTER(C.c_longlong)
vp = C.c_void_p

CPS_OK = 0
STATUS = {0: "CPS_OK", -1: "CPS_ERR_INVALID", -2: "CPS_ERR_CUDA", -3: "CPS_ERR_UNSUPPORTED", -4: "CPS_ERR_NOMEM"}


class CpsError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CpsError(f"{LIB_PATH} is missing: build it with `python -m curvepointslam_b200.build` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.cps_last_error.restype = C.c_char_p
    L.cps_lm_create.argtypes = [C.POINTER(vp), C.c_int]
    L.cps_lm_destroy.argtypes = [vp]
    L.cps_lm_destroy.restype = None
    L.cps_lm_sync.argtypes = [vp]
    host_args = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]
    L.cps_dlevmar_der_batched.argtypes = host_args
    L.cps_dlevmar_dif_batched.argtypes = host_args
    L.cps_dlevmar_batched_pix.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp]
    L.cps_lm_fit_batched_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp,
                                         vp, vp, vp, vp]
    L.cps_lm_set_path.argtypes = [vp, C.c_int]
    L.cps_lm_set_async.argtypes = [vp, C.c_int]
    L.cps_lm_set_arith.argtypes = [vp, C.c_int]
    L.cps_lm_set_profiling.argtypes = [vp, C.c_int]
    L.cps_lm_get_profile.argtypes = [vp, dp, lp, C.c_int]
    L.cps_lm_get_profile_work.argtypes = [vp, vp, vp, C.c_int]
    L.cps_lm_get_solve_handovers.argtypes = [vp, vp]
    L.cps_lm_launch_stats.argtypes = [vp, ip, ip, ip, lp]
    L.cps_fit_curve_init_batched.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.cps_cleanup_and_group_edges_batched.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, vp]
    L.cps_fit_curve_init_batched_dev.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
    L.cps_fit_frames_batched.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, vp]
    fn_t = C.c_void_p
    L.cps_dlevmar_der.argtypes = [fn_t, fn_t, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.cps_dlevmar_dif.argtypes = [fn_t, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.cps_closest_bezier_pt_batched.argtypes = [C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, vp]
    L.cps_measure_fp64_peak.argtypes = [C.c_int, dp, dp]
    L.cps_measure_dmma_peak.argtypes = [C.c_int, dp]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != CPS_OK:
        msg = lib().cps_last_error()
        raise CpsError(f"{what}: {STATUS.get(rc, rc)} ({msg.decode() if msg else ''})")
