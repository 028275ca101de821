# round 2: TZ lane groups for the larger PUs + merge/skip distortion entry: parity, then the stage times
cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_predict.py tests/test_gpu_intra.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2o_pytest.log; tail -5 gpurun_out/r2o_pytest.log
timeout 300 python profiles/tz_split_probe.py 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-encode --no-cpu-baseline --no-full-search > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2o_bench.json')); print('step', round(d['ms_per_step'],3), round(d['value'],2), {k: round(v,3) for k,v in d['stage_ms_per_step'].items()}, 'e2e', round(d['e2e']['value'],2))"
