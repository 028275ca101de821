# round 2: me_fracw.cu parity again + ncu --set full of the group kernel
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fractional or tz_search_and_frac or full_size or pipelined" 2>&1 | tail -25 > gpurun_out/r2j_pytest.log; tail -6 gpurun_out/r2j_pytest.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fracw_group -c 1 -o gpurun_out/r2j_fracw python bench.py --steps 1 --warmup 1 --no-encode --no-cpu-baseline --no-full-search > gpurun_out/r2j_ncu.log 2>&1; tail -2 gpurun_out/r2j_ncu.log | cut -c1-200
