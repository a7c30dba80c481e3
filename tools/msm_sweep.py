#!/usr/bin/env python3
"""Sweep the MSM window width: python tools/msm_sweep.py 18,20,22,24,26  -> best c per size"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpleworks_b200 import _gen  # noqa: E402
from simpleworks_b200.binding import Backend  # noqa: E402

logs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "18,20,22").split(",")]
be = Backend(0)
nmax = 1 << max(logs)
bases = be.bases_from_powers(_gen.g1_generator_jacobian(), _gen.fr_mont(0x5357423230300001), nmax)
g = torch.Generator(device="cuda").manual_seed(1)
out = {}
for lg in logs:
    n = 1 << lg
    s = torch.randint(-2 ** 63, 2 ** 63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    s[:, 3] &= 0x0FFFFFFFFFFFFFFF
    res = {}
    ref = None
    for c in range(max(8, lg - 10), min(22, lg) + 1):
        if n * ((254 + c - 1) // c) >= 2 ** 32:
            continue
        be.set_msm_window_bits(c)
        r = be.msm(bases, s)
        if ref is None:
            ref = r
        assert (r == ref).all(), (lg, c)
        ts = []
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            be.msm(bases, s)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[c] = min(ts)
    best = min(res, key=res.get)
    out[lg] = {"best_c": best, "ms": res}
    print(f"2^{lg}: best c={best} {res[best]:.2f} ms | " + " ".join(f"{c}:{t:.1f}" for c, t in res.items()), flush=True)
be.set_msm_window_bits(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/msm_sweep.json", "w"), indent=1)
