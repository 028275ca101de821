set -e
cd /root/repo
python profiles/prof_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tz_lockstep|tz_search_kernel" -s 2 -c 2 -f -o gpurun_out/prof_r1h_tz python profiles/prof_step.py 2 > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
