set -e
cd /root/repo
python profiles/prof_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"frac_select" -s 2 -c 1 -f -o gpurun_out/prof_r1h_sel python profiles/prof_step.py 2 > gpurun_out/ncu.log 2>&1
