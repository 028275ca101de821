set -e
cd /root/repo
python profiles/prof_fs.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"full_search_packed" -s 2 -c 1 -f -o gpurun_out/prof_r1f_fs python profiles/prof_fs.py > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
