cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-encode --no-cpu-baseline --no-full-search > gpurun_out/r2v_bench.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2v_bench.json')); print('step', round(d['ms_per_step'],3), round(d['value'],2), {k: round(v,3) for k,v in d['stage_ms_per_step'].items()}, 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],3))"
