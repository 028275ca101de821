# round 2: with the encoder binaries + cfg present: the -m gpu suite (MD5 tests run), default bench (encode legs through the broker)
cd /root/repo
nproc
timeout 1500 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -45 > gpurun_out/r2h_pytest.log; tail -4 gpurun_out/r2h_pytest.log
/usr/bin/time -v timeout 1500 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -3 gpurun_out/r2h_bench.err; wc -c gpurun_out/r2h_bench.json
