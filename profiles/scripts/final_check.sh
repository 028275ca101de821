cd /root/repo
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-encode --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 1500 gpurun_out/bench_quick.json | cut -c1-1500
