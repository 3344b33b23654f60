python -m pytest tests/test_frames_gpu.py tests/test_lm_staged_gpu.py tests/test_lm_gpu.py tests/test_lm_dif_staged_gpu.py tests/test_lm_singular_gpu.py -m gpu -x -q 2>&1 | tail -2
for B in 0 1; do
  echo "buckets=$B: $(CPS_LM_BUCKETS=$B python scripts/frames_probe.py 2>&1 | tail -1 | cut -c280-900)"
done
