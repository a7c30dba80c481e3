#!/usr/bin/env python3
"""universal_setup(100000, 25000, 300000) + index + prove + verify as one unit on the GPU (bench.example_bound_gpu),
three times, with the per-call breakdown."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from simpleworks_b200.binding import Backend
be = Backend(0)
for _ in range(3):
    r, proof, vkb = bench.example_bound_gpu(be)
    print(json.dumps({k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in r[k].items()} for k in ("first_call", "steady")}))
