# round 2: the whole -m gpu suite, smoke, a default bench run, the reference arm
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -45 > gpurun_out/r2b_pytest.log; tail -4 gpurun_out/r2b_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -3 gpurun_out/r2b_bench.err; wc -c gpurun_out/r2b_bench.json
