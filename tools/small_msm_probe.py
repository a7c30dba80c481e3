#!/usr/bin/env python3
"""Stage times of prover-sized MSMs on the plain path (no window tables): python tools/small_msm_probe.py [log_n ...]"""
import json
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from simpleworks_b200 import _gen
from simpleworks_b200.binding import Backend
be = Backend(0)
be.profile(True)
for lg in [int(x) for x in (sys.argv[1:] or ["17", "19", "20", "22"])]:
    n = 1 << lg
    bases = be.bases_from_powers(_gen.g1_generator_jacobian(), _gen.fr_mont(bench.BETA_SEED), n)
    dev = torch.from_numpy(bench.synth_scalars_host(n, 7).view(np.int64)).to("cuda:0")
    for _ in range(3):
        be.msm(bases, dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        e0.record(); be.msm(bases, dev); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(lg, "ms", round(min(ts), 3), "M points/s", round(n / min(ts) / 1e3, 1), json.dumps({k: round(v, 3) for k, v in be.last_stages().items()}), flush=True)
    bases.free()
