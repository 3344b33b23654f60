"""One predict + two updates at L landmarks (kernel names and times via ncu launch list)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from curvepointslam_b200 import ekf
L = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
x, G, ci, pi = ekf.make_synthetic_ekf(L, d=3, seed=0)
kf = ekf.KalmanFilter(capacity=x.size)
kf.set_state(x, None, ci, pi)
kf.set_cov_lowrank(G, 1.0 / 64.0, 0.25)
rng = np.random.default_rng(1)
A = np.stack([np.eye(4)] * 4)
for it in range(3):
    z = np.concatenate([rng.normal(0.0, 1.0, 32), x[2:5]])
    kf.PredictKF()
    kf.UpdateNCurvesAndPoints(z, 4, A, [0, 1, 2, 3])
    kf.sync()
    print(it, kf.last_timing())
kf.close()
