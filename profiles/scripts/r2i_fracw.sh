# round 2: the CTU-group fractional kernel (me_fracw.cu): parity, then the stage times beside the per-tile kernels
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fractional or tz_search_and_frac or full_size or pipelined" 2>&1 | tail -25 > gpurun_out/r2i_pytest.log; tail -6 gpurun_out/r2i_pytest.log
for w in 1 0; do HMGPU_FRAC_WIN=$w timeout 300 python bench.py --steps 5 --warmup 3 --no-encode --no-cpu-baseline --no-full-search > gpurun_out/r2i_bench_win$w.json 2> gpurun_out/r2i_bench_win$w.err
python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_win$w.json')); print('frac_win $w', 'step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms_per_step'].items()}, 'roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3))"; done
