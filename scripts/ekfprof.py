"""Small EKF run used as the ncu target: ekfprof.py [landmarks [point measurements]] (default N = 15012, M = 35)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_extras as b

L = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
npm = int(sys.argv[2]) if len(sys.argv) > 2 else 0
print(b.ekf_update_bench(0, L, n_point_meas=npm, reps=2, warm=1))
