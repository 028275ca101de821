cd /root/repo
for c in 0 98304 262144 393216 589824; do
  HMGPU_PIPE_CHUNK=$c python bench.py --steps 5 --warmup 3 --no-encode --no-cpu-baseline --no-full-search 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunk $c', 'value ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'e2e Gcand/s', round(d['e2e']['value'],2))"
done
