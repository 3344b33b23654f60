"""Drop-in by RELINKING: the reference's curveFittingClass.cpp, compiled unchanged, with its levmar call
(`dlevmar_dif(modelfunc, params, final_meas, m, num_meas, 1000, opts, info, NULL, NULL, &funcdata)`, :681) resolved by
libcps_levmar.so instead of liblevmar.a (oracle/Makefile `gpu_relink` -> oracle/_ref/libcps_ref_gpu.so).  The only
addition on the reference's side is cps_register_model(modelfunc, CPS_MODEL_STEREO19).  CurveFittingClass::fit_curve
then runs its packing, t and initial guess on the host as always and the LM iterations on the GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from curvepointslam_b200 import _lib, lm, synth
from oracle_lib import DIF, STEREO19, ORC_DIR, oracle, refcps

pytestmark = pytest.mark.gpu
GPU_SO = os.path.join(ORC_DIR, "_ref", "libcps_ref_gpu.so")
dp = C.POINTER(C.c_double)


@pytest.mark.skipif(not os.path.exists(GPU_SO) or refcps() is None, reason="oracle/_ref/libcps_ref_gpu.so not built")
def test_reference_fit_curve_relinked_against_the_cuda_path():
    oracle()
    _lib.lib()  # libcps.so
    L = C.CDLL(GPU_SO, mode=os.RTLD_NOW | os.RTLD_LOCAL | getattr(os, "RTLD_DEEPBIND", 0))
    L.refh_fit_curve.argtypes = [dp, dp, C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int]
    L.refh_fit_curve.restype = C.c_double
    # an unregistered callback is refused (no CPU path): fit_curve comes back with the initial guess untouched by LM
    b = synth.make_stereo19_batch(12, K=64, seed=91, round_pixels=True)
    assert L.refh_register_gpu_model() == 0
    ref = refcps()  # the same sources linked against the reference's own levmar (CPU)
    rel, erel = [], []
    for f in range(12):
        meas = np.ascontiguousarray(b["meas"][b["off"][f]:b["off"][f + 1]])
        c = b["counts"][f]
        o = np.cumsum([0, 2 * c[0], 2 * c[1], 2 * c[2], 2 * c[3]])
        seg = [np.ascontiguousarray(meas[o[i]:o[i + 1]]) for i in range(4)]
        p = np.zeros(19)
        err = L.refh_fit_curve(p.ctypes.data_as(dp), seg[0].ctypes.data_as(dp), int(c[0]), seg[1].ctypes.data_as(dp), int(c[1]),
                               seg[2].ctypes.data_as(dp), int(c[2]), seg[3].ctypes.data_as(dp), int(c[3]))
        p_cpu, err_cpu = ref.fit_curve(meas, c)
        h = ref.last_fit()
        # the GPU ran dlevmar_dif from the initial guess fit_curve computed on the host: the same call through the
        # product's batched entry gives the same bits (deterministic math), then fit_curve's own post-normalisation
        ctx = lm.LmContext(0)
        g = ctx.fit_batched(STEREO19, DIF, h["p0"][None], meas, np.array([0, meas.size]), c[None], list(h["opts"]))
        ctx.close()
        assert np.array_equal(lm.postfit_normalise(g["p"])[0].view(np.int64), p.view(np.int64))
        assert err == g["info"][0, 1]
        rel.append(np.abs(p - p_cpu).max() / np.abs(p_cpu).max())
        erel.append(abs(err - err_cpu) / err_cpu)
    # against the all-CPU reference (libm model): the distribution two builds of the reference show among themselves
    assert np.median(rel) < 1e-7 and max(rel) < 1e-3 and max(erel) < 1e-6
