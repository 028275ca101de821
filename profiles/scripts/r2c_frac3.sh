# round 2: the fused TMA fractional kernel -- parity first, then the stage time for every ring depth / buffer size / swizzle
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fractional_kernels_agree or frac_only or tz_search_and_frac or full_size or pipelined" 2>&1 | tail -25 > gpurun_out/r2c_pytest.log; tail -5 gpurun_out/r2c_pytest.log
for swz in 1 0; do for v in 0 1 2 3; do
  HMGPU_FRAC3_SWIZZLE=$swz HMGPU_FRAC3_VARIANT=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-encode --no-cpu-baseline --no-full-search > gpurun_out/r2c_bench_s${swz}_v${v}.json 2> gpurun_out/r2c_bench_s${swz}_v${v}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2c_bench_s${swz}_v${v}.json"))
    print("swz $swz variant $v", "ms/step %.3f" % d["ms_per_step"], {k: round(x, 3) for k, x in d["stage_ms_per_step"].items()}, "roofline %.3f" % d["roofline"]["frac"])
except Exception as e:
    print("swz $swz variant $v failed", e)
PY
done; done
HMGPU_FRAC_TMA=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-encode --no-cpu-baseline --no-full-search > gpurun_out/r2c_bench_old.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_old.json')); print('old kernels', d['ms_per_step'], d['stage_ms_per_step'])"
