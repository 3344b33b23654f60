python -m pytest tests/test_frames_gpu.py tests/test_lm_staged_gpu.py tests/test_lm_gpu.py -m gpu -x -q 2>&1 | tail -2
for B in 0 1; do
  CPS_LM_BUCKETS=$B python scripts/frames_probe.py 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()+'\"}' if False else sys.stdin.read()[:0] or '{}')" 2>/dev/null
  echo "buckets=$B: $(CPS_LM_BUCKETS=$B python scripts/frames_probe.py 2>&1 | tail -1 | cut -c1-600)"
done
