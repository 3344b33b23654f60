for P in default 8192 32768 65536 131072; do
  if [ $P = default ]; then unset CPS_LM_STAGED_PIECE; else export CPS_LM_STAGED_PIECE=$P; fi
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('piece $P', 'value', round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3))"
done
