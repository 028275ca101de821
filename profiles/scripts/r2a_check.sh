# round 2, first GPU pass: the whole -m gpu suite (broker tests included), smoke, a default bench run
cd /root/repo
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 > gpurun_out/r2a_pytest.log; tail -5 gpurun_out/r2a_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -3 gpurun_out/r2a_bench.err; wc -c gpurun_out/r2a_bench.json
