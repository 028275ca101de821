cd /root/repo
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tz_ or 1080p" 2>&1 | tail -2
HMGPU_TZ_GROUPS=0 python profiles/tz_ab.py 2>&1 | tail -2 | cut -c1-220
python profiles/tz_ab.py 2>&1 | tail -2 | cut -c1-220
python profiles/tz_ab.py 2>&1 | tail -2 | cut -c1-220
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"tz" -c 6 --csv --log-file gpurun_out/launches_r1n.csv python profiles/prof_step.py 1 > gpurun_out/ncu1.log 2>&1
