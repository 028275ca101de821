cd /root/repo
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-encode --no-cpu-baseline --no-full-search > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-encode --no-cpu-baseline --no-full-search > gpurun_out/ncu_bench.log 2>&1
du -sh gpurun_out
