cd /root/repo
ncu --set full --clock-control none --import-source on -k regex:"tzt_search_kernel" -s 4 -c 1 -f -o gpurun_out/prof_r1n_tzt python profiles/prof_step.py 1 > gpurun_out/ncu_tzt.log 2>&1
tail -2 gpurun_out/ncu_tzt.log
