"""A C++ caller built against include/cps_compat.hpp (the header-only mirror of the reference's KalmanFilter /
CurveFittingClass call surface) and libcps.so.  CPU: it compiles and links.  GPU: it runs and its results equal
the oracle's (EKF within 1e-9, fit bit-exact)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from curvepointslam_b200 import build, synth
from oracle_lib import DIF, STEREO19, oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "compat_smoke")


def compile_exe():
    build.build()
    pkg = os.path.join(ROOT, "curvepointslam_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "compat_smoke.cpp"),
           "-L", pkg, "-lcps", f"-Wl,-rpath,{pkg}", "-o", EXE]
    subprocess.run(cmd, check=True)


def test_cpp_caller_compiles_and_links():
    compile_exe()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_caller_matches_oracle(tmp_path):
    compile_exe()
    g = np.load(os.path.join(ROOT, "tests", "golden", "ekf_curves4.npz"))
    N, nc, n, K = g["P"].shape[0], 6, 4, 48
    b = synth.make_stereo19_batch(1, K=K, seed=31)
    z = np.zeros(8 * n + 3)
    z[:] = np.concatenate([g["z"][:32], g["z"][32:35]])
    A = np.ascontiguousarray(g["A"], dtype=np.float64)
    curve_col = 12 + 8 * np.arange(n)
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("4i", N, nc, n, K))
        for a in (g["x"], g["P"], z, A, b["p0"][0], b["meas"]):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    subprocess.run([EXE, str(inp), str(out)], check=True)
    res = np.fromfile(out, dtype=np.float64)
    x, P, p, err = res[:N], res[N:N + N * N].reshape(N, N), res[N + N * N:N + N * N + 19], res[-1]
    xo, Po = oracle().ekf_predict(g["x"], g["P"])
    xo, Po = oracle().ekf_update_curves_points(xo, Po, z, A, curve_col)
    rel = lambda a, bb: np.abs(a - bb).max() / np.abs(bb).max()
    assert rel(x, xo) < 1e-9 and rel(P, Po) < 1e-9
    # the reference writes opts[4] = LM_DIFF_DELTA*1E1 (curveFittingClass.cpp:662), which is NOT the double 1e-5
    # like the reference, the mirror's fit_curve ignores params on entry: the initial guess is the stereo triangulation
    # + plane fit of curveFittingClass.cpp:516-651
    import ctypes as C
    dp = C.POINTER(C.c_double)
    p0 = np.zeros((1, 19))
    cnt = np.ascontiguousarray(b["counts"][0], dtype=np.int32)
    m0 = np.ascontiguousarray(b["meas"])
    oracle().lib.orc_stereo19_init_guess(m0.ctypes.data_as(dp), cnt.ctypes.data_as(C.POINTER(C.c_int)), p0[0].ctypes.data_as(dp), 0, 1)
    fit = oracle().fit_batch(STEREO19, DIF, p0, b["meas"], b["off"], b["counts"], [1e-10, 1e-10, 1e-10, 1e-20, 1e-6 * 1e1])
    exp = fit["p"][0].copy()
    oracle().lib.orc_stereo19_postfit(exp.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.array_equal(p.view(np.int64), exp.view(np.int64)) and err == fit["info"][0, 1]
