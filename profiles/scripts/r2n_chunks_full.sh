# round 2: end-to-end chunk sweep, then the encode legs at BASELINE size (bench.py --full)
cd /root/repo
for c in 0 589320 392880 196440; do HMGPU_PIPE_CHUNK=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-encode --no-cpu-baseline --no-full-search > gpurun_out/r2n_chunk$c.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2n_chunk$c.json')); print('pipe_chunk $c', 'device', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), round(d['e2e']['value'],2))"; done
S=$(date +%s); timeout 3300 python bench.py --full > gpurun_out/r2n_bench_full.json 2> gpurun_out/r2n_bench_full.err; echo "bench --full rc $? in $(( $(date +%s) - S )) s"; tail -2 gpurun_out/r2n_bench_full.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench_full.json'))
for k in ('encode','encode_segments','encode_shared_gpu'):
    print(k, json.dumps(d.get(k))[:1500])
PY
