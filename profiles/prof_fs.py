"""One device-resident step of the full-search workload (BASELINE configs[0] as a batch), for ncu.
usage: python profiles/prof_fs.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

print(bench.full_search_leg(0, 1, 1.0, 1.0))
