set -e
cd /root/repo
python profiles/prof_step.py 1 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"tz" -c 30 --csv --log-file gpurun_out/launches_r1n.csv python profiles/prof_step.py 1 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tzt_search_kernel" -s 3 -c 1 -f -o gpurun_out/prof_r1n_tzt python profiles/prof_step.py 1 > gpurun_out/ncu_tzt.log 2>&1
