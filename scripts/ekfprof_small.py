"""Small EKF run used as the ncu target for the front end (N = 3012, M = 35)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_extras as b

print(b.ekf_update_bench(0, 1000, reps=3, warm=1))
