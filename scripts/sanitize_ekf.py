"""The EKF kernels of the second half of round 2 at a size where the streaming path runs (N = 3 012): predict, updates
with M = 35 (two-group kernel, 4-deep ring), M = 65 (3-deep ring), M = 80 (five front kernels + single-group ring),
read-back below the diagonal tiles, gate, augmentation.  Meant to be run under `compute-sanitizer --tool memcheck`
or `--tool racecheck` (one tool per gpurun call)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curvepointslam_b200 import ekf

x, G, ci, pi = ekf.make_synthetic_ekf(1000, d=3, seed=3)
kf = ekf.KalmanFilter(capacity=x.size + 64)
kf.set_state(x, None, ci, pi)
kf.set_cov_lowrank(G, 1.0 / 64.0, 0.25)
rng = np.random.default_rng(2)
A = np.stack([np.eye(4)] * 4)
for n_pts in (0, 10, 15):
    z = np.concatenate([rng.normal(0, 1.0, 32), x[2:5]])
    pts = rng.choice(len(pi), n_pts, replace=False) if n_pts else None
    pm = rng.normal(0, 3.0, 3 * n_pts) if n_pts else None
    kf.PredictKF()
    kf.UpdateNCurvesAndPoints(z, 4, A, [0, 1, 2, 3], pm, pts, n_pts)
    kf.sync()
    print("update M =", 35 + 3 * n_pts, "ok")
print("gate", kf.gate(rng.normal(0, 30.0, (50, 3)), 11.345)[0][:5])
kf.AddNewPoints(rng.normal(0, 5.0, 6), 2)
kf.PredictKF()
kf.UpdatePoints(rng.normal(0, 3.0, 6), [len(pi), len(pi) + 1], 2)
P = kf.getCov()
print("symmetric", np.array_equal(P, P.T), "finite", np.isfinite(P).all())
kf.close()
