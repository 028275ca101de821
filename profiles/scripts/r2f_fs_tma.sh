cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2f_pytest.log; tail -4 gpurun_out/r2f_pytest.log
for t in 1 0; do HMGPU_FS_TMA=$t timeout 300 python bench.py --steps 6 --warmup 3 --no-encode --no-cpu-baseline > gpurun_out/r2f_bench_tma$t.json 2> gpurun_out/r2f_bench_tma$t.err
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_tma$t.json')); print('fs_tma $t', 'step', round(d['ms_per_step'],3), d['stage_ms_per_step'], 'full search ms', round(d['full_search']['ms_per_step'],3), round(d['full_search']['gcand_per_s'],1))"; done
