cd /root/repo
rm -f gpurun_out/*.ncu-rep
# one step of the bench workload under ncu --set full: the TZ stage (17 launches) and the fractional stage (4 launches) of the second step
python profiles/prof_step.py 2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tz" -s 17 -c 17 -f -o /tmp/prof_r1n_tz python profiles/prof_step.py 2 > gpurun_out/ncu_tz.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"frac2_dist" -s 4 -c 4 -f -o /tmp/prof_r1n_frac python profiles/prof_step.py 2 > gpurun_out/ncu_frac.log 2>&1
python profiles/summarize_ncu.py /tmp/prof_r1n_tz.ncu-rep gpurun_out/r1n_ncu_tz_stage.csv
python profiles/summarize_ncu.py /tmp/prof_r1n_frac.ncu-rep gpurun_out/r1n_ncu_frac2_dist.csv
# source-level hot spots: launch 1 of the TZ capture = the 4x8 shape kernel, 3 = 8x8; the last two = warp-per-job kernel; fractional 0 = 8x8 tiles, 1 = 4x4
python profiles/ncu_lines.py /tmp/prof_r1n_tz.ncu-rep 1 30 > gpurun_out/r1n_ncu_tz_thread_4x8_lines.txt 2>&1
python profiles/ncu_lines.py /tmp/prof_r1n_tz.ncu-rep 3 30 > gpurun_out/r1n_ncu_tz_thread_8x8_lines.txt 2>&1
python profiles/ncu_lines.py /tmp/prof_r1n_tz.ncu-rep 15 30 > gpurun_out/r1n_ncu_tz_warp_lines.txt 2>&1
python profiles/ncu_lines.py /tmp/prof_r1n_frac.ncu-rep 0 30 > gpurun_out/r1n_ncu_frac2_dist8_lines.txt 2>&1
python profiles/ncu_lines.py /tmp/prof_r1n_frac.ncu-rep 1 30 > gpurun_out/r1n_ncu_frac2_dist4_lines.txt 2>&1
ncu -i /tmp/prof_r1n_tz.ncu-rep --page details --launch-skip 3 --launch-count 1 > gpurun_out/r1n_ncu_tz_thread_8x8_details.txt 2>&1
ls -la /tmp/*.ncu-rep; du -sh gpurun_out
