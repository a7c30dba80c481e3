#!/usr/bin/env python3
"""setup / index / prove / verify timings of the mul-chain circuit at 2^k constraints (bench.marlin_gpu_run)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from simpleworks_b200.binding import Backend
be = Backend(0)
for lg in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "16,18,20").split(",")]:
    r, _ = bench.marlin_gpu_run(be, lg, 3)
    print(lg, {k: round(r[k], 4) for k in ("setup_s", "index_s", "prove_s", "prove_s_plain_msm_path", "verify_s")}, r["verified"])
    print("   prove phases", json.dumps(r["prove_phases_ms"]))
    print("   index phases", json.dumps(r["index_phases_ms"]))
