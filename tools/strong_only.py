"""bench.py's strong-scaling sweep (BASELINE configs[4]) alone, on one GPU.  usage: python tools/strong_only.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
import bench

pkg = entry.load_package()
cfg, pat, _ = bench.workload(pkg)
r = bench.strong_scaling(pkg, pat, 0, 1, 0, torch, None)
print({k: r[k] for k in ("qps_device_timed", "device_ms_max_over_ranks", "iters_mean")})
