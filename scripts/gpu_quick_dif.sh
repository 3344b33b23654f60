cd $GRAFT_REPO_ROOT
make -s -C oracle 2>&1 | tail -2
python -m pytest tests/test_lm_gpu.py -x -q 2>&1 | tail -3
python bench.py --variant dif --fits 32768 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/b2.json 2>gpurun_out/b2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/b2.json').read().strip().splitlines()[-1])
print('dif value %.4g e2e %.4g kernel_ms %.3f frac %.4f'%(d['value'],d['e2e']['value'],d['roofline']['kernel_ms'],d['roofline']['frac']))
PY
