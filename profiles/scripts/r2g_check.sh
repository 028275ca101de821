# round 2, after the container restart: the whole -m gpu suite, smoke, default bench, the reference arm, ncu launch list
cd /root/repo
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc
timeout 1500 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -45 > gpurun_out/r2g_pytest.log; tail -4 gpurun_out/r2g_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -3 gpurun_out/r2g_bench.err; wc -c gpurun_out/r2g_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/r2g_bench_ref.json 2> gpurun_out/r2g_bench_ref.err; wc -c gpurun_out/r2g_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-encode --no-cpu-baseline > gpurun_out/r2g_ncu.log 2>&1; tail -2 gpurun_out/r2g_ncu.log | cut -c1-300
