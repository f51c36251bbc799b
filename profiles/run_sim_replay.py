"""Driver of the ncu capture of sim_kernel: one swb_sim_replay of 1184 scenarios x 643 rounds (same as bench.py's leg)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

print(json.dumps(bench.bench_sim_loop(0)))
