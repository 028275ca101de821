cd /root/repo
for e in 8 12 16 24; do HMGPU_PIPE_EDGE=$e timeout 300 python bench.py --steps 5 --warmup 3 --no-encode --no-cpu-baseline --no-full-search > gpurun_out/r2p_edge$e.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2p_edge$e.json')); print('pipe_edge $e', 'device', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), round(d['e2e']['value'],2))"; done
