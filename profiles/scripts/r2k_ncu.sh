cd /root/repo
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fracw_group -c 1 -o gpurun_out/r2k_fracw python bench.py --steps 1 --warmup 1 --no-encode --no-cpu-baseline --no-full-search > gpurun_out/r2k_ncu.log 2>&1; tail -2 gpurun_out/r2k_ncu.log | cut -c1-200
