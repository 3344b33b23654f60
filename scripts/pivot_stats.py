"""How often does the reference's scaled partial pivoting (Axb_core.c:1044-1132) leave the diagonal on J^T J + mu I of
the Stereo19 model?  Decides whether a block-arrow storage of the solve's work matrix is worth building."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from curvepointslam_b200 import synth
from oracle_lib import oracle, STEREO19, DER

class FD(C.Structure):
    _fields_ = [("t", C.POINTER(C.c_double)), ("count", C.c_int * 4), ("math_mode", C.c_int)]

def pivots(A):
    m = A.shape[0]
    a = A.copy()
    scale = 1.0 / np.abs(a).max(axis=1)
    piv = -1
    out = []
    for j in range(m):
        big = 0.0
        for i in range(j, m):
            merit = scale[i] * abs(a[i, j])
            if merit >= big:
                big, piv = merit, i
        out.append(piv)
        if piv != j:
            a[[piv, j]] = a[[j, piv]]
            scale[piv] = scale[j]
        if a[j, j] == 0.0:
            a[j, j] = np.finfo(float).eps
        if j != m - 1:
            a[j + 1:, j] /= a[j, j]
            a[j + 1:, j + 1:] -= np.outer(a[j + 1:, j], a[j, j + 1:])
    return out

o = oracle()
B = 300
b = synth.make_stereo19_batch(B, K=64, seed=5)
fit = o.fit_batch(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], [1e-3, 1e-15, 1e-15, 1e-20, 1e-6])
dp = C.POINTER(C.c_double)
stats = {}
for which, P in (("p0", b["p0"]), ("fitted", fit["p"])):
    for mu_rel in (1e-3, 1e-1, 10.0):
        off_diag = 0; total = 0; to_tail = 0; cross = 0; fits_bad = 0
        for f in range(B):
            meas = np.ascontiguousarray(b["meas"][b["off"][f]:b["off"][f + 1]])
            cnt = b["counts"][f]
            n = meas.size
            t = np.zeros(n // 2)
            s = 0
            for sg in range(4):
                seg = np.ascontiguousarray(meas[2 * s:2 * (s + cnt[sg])])
                o.lib.orc_segment_t(seg.ctypes.data_as(dp), int(cnt[sg]), t[s:].ctypes.data_as(dp))
                s += cnt[sg]
            fd = FD(t.ctypes.data_as(dp), (C.c_int * 4)(*[int(c) for c in cnt]), 0)
            J = np.zeros((n, 19))
            p = np.ascontiguousarray(P[f])
            o.lib.orc_stereo19_jac(p.ctypes.data_as(dp), J.ctypes.data_as(dp), 19, n, C.byref(fd))
            A = J.T @ J
            A = A + mu_rel * A.diagonal().max() * np.eye(19)
            pv = pivots(A)
            bad = False
            for j, r in enumerate(pv):
                total += 1
                if r != j:
                    off_diag += 1
                    if j < 16 and r >= 16: to_tail += 1; bad = True
                    if j < 8 and 8 <= r < 16: cross += 1; bad = True
            fits_bad += bad
        print(which, mu_rel, "off-diagonal pivots %.3f" % (off_diag / total), "curve col -> theta/phi/z row", to_tail, "curve1 col -> curve2 row", cross, "fits leaving the block structure %.3f" % (fits_bad / B))
