cd $GRAFT_REPO_ROOT
make -s -C oracle 2>&1 | tail -2
python -m pytest tests/test_ekf_gpu.py -x -q 2>&1 | tail -5
python - <<'PY'
from curvepointslam_b200 import bench_extras as b
for L in (1000, 5000, 20000):
    r = b.ekf_update_bench(0, L)
    print('PIPE', L, 'update_ms %.4f cov_ms %.4f frac %.3f' % (r['update_ms'], r['cov_kernel_ms'], r['roofline']['frac']))
PY
CPS_EKF_COV_NO_PIPE=1 python - <<'PY'
from curvepointslam_b200 import bench_extras as b
for L in (5000, 20000):
    r = b.ekf_update_bench(0, L)
    print('DMMA-nopipe', L, 'update_ms %.4f cov_ms %.4f frac %.3f' % (r['update_ms'], r['cov_kernel_ms'], r['roofline']['frac']))
PY
