# round 2: whole -m gpu suite, smoke, default bench, reference arm, ncu launch list of the bench command
cd /root/repo
S=$(date +%s); timeout 2400 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -30 > gpurun_out/r2t_pytest.log; tail -3 gpurun_out/r2t_pytest.log; echo "gpu suite $(( $(date +%s) - S )) s"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
S=$(date +%s); timeout 1500 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc $? in $(( $(date +%s) - S )) s"; tail -2 gpurun_out/r2t_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2t_bench_ref.json 2> gpurun_out/r2t_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-encode --no-cpu-baseline --no-full-search > gpurun_out/r2t_ncu.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2t_bench.json'))
print('value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), {k: round(v,3) for k,v in d['stage_ms_per_step'].items()})
print('roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'longest', d['roofline']['longest_stage']['stage'], round(d['roofline']['longest_stage']['frac_int32'],3), 'me_step', round(d['roofline_int32']['me_step']['frac'],3))
for k in ('encode','encode_segments','encode_shared_gpu'):
    print(k, json.dumps(d.get(k))[:600])
PY
