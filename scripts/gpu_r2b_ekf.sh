#!/bin/bash
# round 2, second half: EKF front end (register-resident Gauss-Jordan, conflict-free tables) and the two-stage TMA
# covariance kernel for M <= 80
mkdir -p gpurun_out
python -m pytest tests/test_ekf_gpu.py tests/test_baseline_configs_gpu.py -m gpu -x -q > gpurun_out/r2b_gpu_tests_ekf.txt 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2b_gpu_tests_ekf.txt
python scripts/ekf_bench_probe.py > gpurun_out/r2b_ekf.json 2> gpurun_out/r2b_ekf.err; echo "ekf rc=$?"
cat gpurun_out/r2b_ekf.json
for cfg in "1000 0" "5000 10"; do
  tag=$(echo $cfg | tr ' ' '_')
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_ekf_launches_$tag.csv \
    python scripts/ekfprof.py $cfg > gpurun_out/r2b_ncu_ekf_$tag.log 2>&1; echo "ncu $tag rc=$?"
done
