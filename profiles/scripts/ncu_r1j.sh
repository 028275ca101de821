set -e
cd /root/repo
python profiles/prof_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1j.csv python profiles/prof_step.py 2 > gpurun_out/ncu1.log 2>&1
python profiles/prof_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"frac2_dist_kernel" -s 4 -c 2 -f -o gpurun_out/prof_r1j_frac python profiles/prof_step.py 2 > gpurun_out/ncu2.log 2>&1
