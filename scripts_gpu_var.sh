cd $GRAFT_REPO_ROOT
for c in 4 6 8 10 12; do
  CPS_LM_MAX_WARPS_PER_SM=$c python bench.py --variant dif --fits 32768 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bv.json 2>gpurun_out/bv.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bv.json').read().strip().splitlines()[-1])
    print('cap $c', 'dif value %.4g kernel_ms %.3f'%(d['value'],d['roofline']['kernel_ms']))
except Exception as e:
    print('cap $c', 'ERR', open('gpurun_out/bv.err').read()[-300:])
PY
done
for c in 6 8 10; do
  CPS_LM_MAX_WARPS_PER_SM=$c python bench.py --fits 65536 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bv.json 2>gpurun_out/bv.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bv.json').read().strip().splitlines()[-1])
print('cap $c', 'der value %.4g kernel_ms %.3f'%(d['value'],d['roofline']['kernel_ms']))
PY
done
