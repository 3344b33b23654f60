for cfg in "1000 0" "5000 10"; do
  tag=$(echo $cfg | tr ' ' '_')
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2b_ekf_launches_$tag.csv python scripts/ekfprof.py $cfg > gpurun_out/r2b_ncu_ekf_$tag.log 2>&1
done
python -m pytest tests/test_ekf_gpu.py -m gpu -x -q 2>&1 | tail -2
