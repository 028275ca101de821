cd /root/repo
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "frac or tz_search" 2>&1 | tail -2
python profiles/tz_ab.py 2>&1 | tail -2 | cut -c1-220
