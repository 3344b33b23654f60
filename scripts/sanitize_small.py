"""One small invocation of every kernel, meant to be run under `compute-sanitizer --tool memcheck` (one tool per
gpurun call, see B200_PROFILING.md)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curvepointslam_b200 import ekf, lm, synth

ctx = lm.LmContext(0)
rng = np.random.default_rng(1)
counts = rng.integers(5, 102, size=(12, 4)).astype(np.int32)
counts[0] = (3, 2, 3, 2)
counts[1] = (101, 101, 101, 101)
b = synth.make_stereo19_batch(12, seed=3, counts=counts)
for var in (lm.DER, lm.DIF):
    r = ctx.fit_batched(lm.STEREO19, var, b["p0"], b["meas"], b["off"], b["counts"])
    print("stereo19", var, r["ret"][:4])
# the staged shape (phase kernels) on the same ragged batch plus a mixed-block fit, and the m = 8 thread-per-fit kernel
sctx = lm.LmContext(0)
sctx.set_path(lm.PATH_STAGED)
counts2 = counts.copy()
counts2[2] = (21, 3, 22, 30)
counts2[3] = (7, 7, 6, 7)
b2 = synth.make_stereo19_batch(12, seed=3, counts=counts2)
r = sctx.fit_batched(lm.STEREO19, lm.DER, b2["p0"], b2["meas"], b2["off"], b2["counts"])
print("stereo19 staged", r["ret"][:4])
for K in (33, 64, 70):
    b8s = synth.make_image8_batch(40, K=K, seed=4)
    r = sctx.fit_batched(lm.IMAGE8, lm.DER, b8s["p0"], b8s["meas"], b8s["off"], b8s["counts"])
    print("image8 thread-per-fit", K, r["ret"][:4])
sctx.close()
b8 = synth.make_image8_batch(40, K=33, seed=4)
for var in (lm.DER, lm.DIF):
    r = ctx.fit_batched(lm.IMAGE8, var, b8["p0"], b8["meas"], b8["off"], b8["counts"])
    print("image8", var, r["ret"][:4])
print("init", ctx.init_guess(b["meas"], b["off"], b["counts"])[1, 16:])
pts, so, tr = synth.make_edge_frames(20, seed=5)
print("edges", ctx.cleanup_and_group_edges(pts, so, tr)["valid"].sum())
print("closest", lm.closest_bezier_pt(rng.uniform(0, 100, (50, 8)), rng.uniform(0, 100, (50, 2)))[1][:3])
x, G, ci, pi = ekf.make_synthetic_ekf(60, d=3, seed=6)
N = x.size
P = G @ G.T / 64 * 0.05 + 0.25 * np.eye(N)
kf = ekf.KalmanFilter(capacity=N + 40)
kf.set_state(x, 0.5 * (P + P.T), ci, pi)
kf.PredictKF()
kf.UpdateNCurvesAndPoints(np.concatenate([rng.normal(0, 1, 32), x[2:5]]), 4, np.stack([np.eye(4)] * 4), [0, 1, 2, 3],
                          rng.normal(0, 3, 6), [2, 5], 2)
kf.UpdatePoints(rng.normal(0, 3, 9), [1, 3, 7], 3)
kf.AddNewCurve(rng.normal(3, 1, 8), np.eye(4))
kf.AddNewPoints(rng.normal(0, 3, 6), 2)
print("ekf", kf.N, kf.getState()[:3], kf.gate(rng.uniform(-50, 50, (40, 3)))[0][:5])
kf2 = ekf.KalmanFilter(capacity=64)
kf2.AddFirstStates(rng.normal(0, 1, 19))
print("first", kf2.N)
print("sanitize_small done")
