# round 2: default bench with the encode legs (through the broker)
cd /root/repo
nproc
S=$(date +%s); timeout 1500 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc $? in $(( $(date +%s) - S )) s"; tail -3 gpurun_out/r2h_bench.err; wc -c gpurun_out/r2h_bench.json
